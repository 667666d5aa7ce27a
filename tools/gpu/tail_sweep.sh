# tail-kernel threshold / start sweep on one scene: bash tools/gpu/tail_sweep.sh terrain "1 2 3" "49152 98304"
SCENE=${1:-terrain}
STARTS=${2:-"1 2 3"}
THRS=${3:-"49152 98304 196608 393216"}
for st in $STARTS; do for th in $THRS; do
  echo -n "${TAG}start$st/thr$th "; HL_TAIL_START=$st HL_TAIL_THRESHOLD=$th python tools/frame_time.py $SCENE
done; done
echo -n "${TAG}notail "; HL_TAIL_THRESHOLD=0 python tools/frame_time.py $SCENE
