// f64 instantiations of the shifted second pass of the pair kernel (pair_kernel<..., SHIFT = true>): its own
// translation unit so that it builds in parallel with pair_f64.cu.
#define PBN_T double
#define PBN_SHIFT_LAUNCH_NAME launch_pair_shift_f64
#define PBN_SHIFT_WARM_NAME warm_pair_shift_f64
#include "pair_launch.inl"
