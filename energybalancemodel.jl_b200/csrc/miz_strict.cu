// miz_strict.cu -- literal-arithmetic flavour of the MIZ kernel: compiled with -fmad=false (no FMA contraction),
// reference operation order, IEEE division, serial Thomas solve in the oracle's order.  Slow; exists for parity
// debugging (ebm_options_t.strict) and for the one-step entry point ebm_miz_step.

#include "miz_literal.cuh"

int ebm_launch_miz_strict(const MizKArgs& a, cudaStream_t stream) { return miz_launch_any(a, stream); }
