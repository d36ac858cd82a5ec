// f32 pair kernel with group skipping inside the units (pass B of tile skipping): own translation unit, builds in parallel.
#define PBN_T float
#define PBN_GSKIP_LAUNCH_NAME launch_pair_gskip_f32
#define PBN_GSKIP_WARM_NAME warm_pair_gskip_f32
#include "pair_launch.inl"
