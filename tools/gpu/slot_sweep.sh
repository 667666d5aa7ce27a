# frames-in-flight x tail-kernel sweep: bash tools/gpu/slot_sweep.sh terrain "4 6 8" "3 4 5" "32768 98304"
SCENE=${1:-terrain}
for n in ${2:-"4 6 8"}; do for st in ${3:-"3 4"}; do for th in ${4:-"32768 98304"}; do
  echo -n "s$n/start$st/thr$th "; HL_SLOTS=$n HL_TAIL_START=$st HL_TAIL_THRESHOLD=$th python tools/frame_time.py $SCENE
done; done; echo -n "s$n/notail "; HL_SLOTS=$n HL_TAIL_THRESHOLD=0 python tools/frame_time.py $SCENE; done
