/* TEST INFRASTRUCTURE ONLY.  Stands in for the reference's pub/mp3enc.h when oracle/Makefile compiles the
 * reference's UNMODIFIED command line (hmp3/src/test/tomp3.cpp) against the GPU library: `CMp3Enc Encode;`
 * (tomp3.cpp:664) becomes the C-ABI adapter of include/cmp3enc_gpu.h.  Nothing of the reference's encoder is
 * linked into that binary (only its WAV parser pcmhpm.c and Xing writer xhead.c, which tomp3.cpp calls itself). */
#ifndef HMP3_SHIM_MP3ENC_H_
#define HMP3_SHIM_MP3ENC_H_
#include <assert.h>
#include "encapp.h"
#include "hxtypes.h" /* min/max macros the CLI expects from pub/mp3enc.h:56 */
#include "cmp3enc_gpu.h"
typedef CMp3EncGpu CMp3Enc;
#endif
