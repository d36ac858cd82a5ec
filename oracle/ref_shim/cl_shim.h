// cl_shim.h — just enough of the OpenCL C 1.2 execution model to compile the reference's
// kernel source (pybnesian/kde/opencl_kernels/KDE.cl.src, expanded by its own
// conv_template.py) as C++ and run it on the host.  TEST INFRASTRUCTURE ONLY (oracle/_ref).
//
// A kernel launch is emulated by ref_driver.cpp: barrier-free kernels run one work-item after
// another; kernels that use barrier() run one OpenMP thread per work-item of a work-group, with
// barrier() mapped to an OpenMP barrier.
#pragma once
#include <math.h>
#include <stddef.h>

#define __kernel
#define __global
#define __private
#define __local
#define __constant const
#define restrict __restrict__
#define CLK_LOCAL_MEM_FENCE 0
#define barrier(flags) _Pragma("omp barrier")
#ifndef M_SQRT1_2_F
#define M_SQRT1_2_F 0.70710678118654752440f
#endif
#ifndef M_LN2_F
#define M_LN2_F 0.69314718055994530942f
#endif

typedef unsigned int uint;

struct WorkItem {
    size_t global_id[3], global_size[3], local_id[3], local_size[3], group_id[3], num_groups[3];
};
extern thread_local WorkItem g_wi;

static inline size_t get_global_id(uint d) { return g_wi.global_id[d]; }
static inline size_t get_global_size(uint d) { return g_wi.global_size[d]; }
static inline size_t get_local_id(uint d) { return g_wi.local_id[d]; }
static inline size_t get_local_size(uint d) { return g_wi.local_size[d]; }
static inline size_t get_group_id(uint d) { return g_wi.group_id[d]; }
static inline size_t get_num_groups(uint d) { return g_wi.num_groups[d]; }

static inline float max(float a, float b) { return a < b ? b : a; }
static inline double max(double a, double b) { return a < b ? b : a; }
