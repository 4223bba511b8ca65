// miz_kernel.cu -- placeholder until the MIZ kernel lands (next commit).
#include "ebm_internal.cuh"
int ebm_launch_miz(const MizKArgs&, int, cudaStream_t) { ebm_set_error("MIZ kernel not built yet"); return EBM_ERR_UNSUPPORTED; }
int ebm_launch_miz_single_step(const EbmGridTables&, const double*, int, double, double, int, double*, double*, double*, double*,
                               double*, double*, double*, int*, cudaStream_t) { ebm_set_error("MIZ kernel not built yet"); return EBM_ERR_UNSUPPORTED; }
