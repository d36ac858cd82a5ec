// f32 instantiations of the pair kernel (FP32 FMA pipe + MUFU ex2.approx).
#define PBN_T float
#define PBN_LAUNCH_NAME launch_pair_f32
#define PBN_TILE_NAME pair_tile_f32
#define PBN_TB_NAME pair_tb_f32
#define PBN_TB_FOR_NAME pair_tb_for_f32
#define PBN_TB_CDF_NAME pair_tb_cdf_f32
#define PBN_CTAS_NAME pair_ctas_per_sm_f32
#define PBN_WARM_NAME warm_pair_f32
#define PBN_CDF_LAUNCH_NAME launch_cdf_f32
#include "pair_launch.inl"
