// Host-side control resolution and table generation (see enc_init.cpp).
#pragma once
#include "../../include/hmp3_b200.h"
#include "enc_tables.h"

namespace hmp3 {
void control_defaults(hmp3_control *ec);
int control_apply_option(hmp3_control *ec, const char *opt);
// Returns bytes_in (nchan*4*1152) or 0 when the control block is rejected.
int build_tables(const hmp3_control *ec, EncTables *T, int *unsupported);
// The configuration-independent polyphase tables (window fold coefficients [32][8] x 2, DCT twiddles [31]): what
// build_tables puts into EncTables::polyA / polyB / dct32, for the kernels' constant memory.
void fixed_polyphase_tables(float *polyA, float *polyB, float *dct32);
}  // namespace hmp3
