// f64 instantiations of the pair kernel (FP64 FMA pipe + table/polynomial exp2).
#define PBN_T double
#define PBN_LAUNCH_NAME launch_pair_f64
#define PBN_TILE_NAME pair_tile_f64
#define PBN_TB_NAME pair_tb_f64
#define PBN_TB_FOR_NAME pair_tb_for_f64
#define PBN_TB_CDF_NAME pair_tb_cdf_f64
#define PBN_CTAS_NAME pair_ctas_per_sm_f64
#define PBN_WARM_NAME warm_pair_f64
#define PBN_CDF_LAUNCH_NAME launch_cdf_f64
#include "pair_launch.inl"
