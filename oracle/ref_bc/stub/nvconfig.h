// ORACLE — TEST INFRASTRUCTURE ONLY: nvidia-texture-tools generates nvconfig.h with cmake; the decoders need none of its switches.
#ifndef NV_CONFIG
#define NV_CONFIG
#define NV_HAVE_UNISTD_H
#define NV_HAVE_STDARG_H
#endif
