
    #include <stdio.h>
    #include "gmd_b200.h"
    int main(void){ printf("%zu %zu %zu %zu %zu\n", sizeof(gmd_hdr_params), sizeof(gmd_sched_params), sizeof(gmd_gemm_params),
                            sizeof(gmd_conv_params), sizeof(gmd_attn_params)); return 0; }