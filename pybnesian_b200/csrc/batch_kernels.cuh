// batch_kernels.cuh — kernels shared by the multi-job launches (cv.cu: all (candidate, fold) jobs of a score
// batch; segmented.cu: all discrete-configuration slices of a hybrid factor).  Included inside an anonymous
// namespace by each translation unit.
#pragma once

constexpr int kMaxFast = pbn::kMaxFastD;  // the pair kernel is instantiated for d = 1..10

__device__ __forceinline__ double block_sum(double v, double* sh) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double r = 0;
    if (threadIdx.x == 0) {
        int nw = (blockDim.x + 31) >> 5;
        for (int i = 0; i < nw; ++i) r += sh[i];
    }
    return r;  // valid in thread 0
}

// ---- batched finalize / fallback / per-job sums ----------------------------------------------------
struct FinJob {
    double lognorm_joint, lognorm_marg;
    long long out_off;  // first element of this job in the logl buffer
};

__global__ void finalize_batch_kernel(const PairJob* __restrict__ jobs, const FinJob* __restrict__ fin, long long upb, int tb,
                                      int ckde, double thresh, double* __restrict__ out, int2* __restrict__ flagged,
                                      int* __restrict__ n_flagged) {
    const PairJob jb = jobs[blockIdx.y];
    const FinJob fj = fin[blockIdx.y];
    for (long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x; row < jb.m; row += (long long)gridDim.x * blockDim.x) {
        long long tt = row / tb;
        long long ustart = jb.unit_begin + tt * jb.n_train_tiles;
        int first = (int)(ustart / upb);
        int last = (int)((ustart + jb.n_train_tiles - 1) / upb);
        int ns = last - first + 1;
        double sj = 0, sm = 0;
        for (int s = 0; s < ns; ++s) {
            sj += jb.part[(long long)s * jb.m_pad + row];
            if (ckde) sm += jb.part[((long long)jb.slots + s) * jb.m_pad + row];
        }
        bool bad = !(sj >= thresh) || (ckde && !(sm >= thresh));
        if (sj != sj || (ckde && sm != sm)) bad = false;  // NaN inputs propagate
        if (bad) {
            int slot = atomicAdd(n_flagged, 1);
            flagged[slot] = make_int2((int)blockIdx.y, (int)row);
            continue;
        }
        double v = fj.lognorm_joint + log(sj);
        if (ckde) v -= fj.lognorm_marg + log(sm);
        out[fj.out_off + row] = v;
    }
}

// exact max-shifted evaluation of the flagged (job, row) entries, one CTA each (cf. row_kernel in runtime.cu)
template <typename T>
__global__ void row_batch_kernel(const PairJob* __restrict__ jobs, const FinJob* __restrict__ fin, int d, int ckde, double u2,
                                 const int2* __restrict__ flagged, const int* __restrict__ n_flagged, double* __restrict__ out) {
    __shared__ double sh[32];
    __shared__ double bc[2];
    __shared__ double yt[kMaxFast];
    const int cnt = *n_flagged;
    for (int f = blockIdx.x; f < cnt; f += gridDim.x) {
        const int2 fr = flagged[f];
        const PairJob jb = jobs[fr.x];
        const FinJob fj = fin[fr.x];
        const T* tr = static_cast<const T*>(jb.train);
        const T* te = static_cast<const T*>(jb.test);
        const long long row = fr.y;
        __syncthreads();
        if (threadIdx.x < d) yt[threadIdx.x] = static_cast<double>(te[row * d + threadIdx.x]);
        __syncthreads();
        double mnj = INFINITY, mnm = INFINITY;
        for (long long i = threadIdx.x; i < jb.n_train; i += blockDim.x) {
            double s = 0, sm = 0;
            for (int c = 0; c < d; ++c) {
                double dl = yt[c] - static_cast<double>(tr[i * d + c]);
                s = fma(dl, dl, s);
                if (c == d - 2) sm = s;
            }
            mnj = fmin(mnj, s);
            mnm = fmin(mnm, sm);
        }
        for (int o = 16; o > 0; o >>= 1) {
            mnj = fmin(mnj, __shfl_down_sync(0xffffffffu, mnj, o));
            mnm = fmin(mnm, __shfl_down_sync(0xffffffffu, mnm, o));
        }
        __syncthreads();
        if ((threadIdx.x & 31) == 0) { sh[threadIdx.x >> 5] = mnj; sh[16 + (threadIdx.x >> 5)] = mnm; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double a = INFINITY, b = INFINITY;
            for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a = fmin(a, sh[i]); b = fmin(b, sh[16 + i]); }
            bc[0] = a; bc[1] = b;
        }
        __syncthreads();
        mnj = bc[0]; mnm = bc[1];
        double aj = 0, am = 0;
        for (long long i = threadIdx.x; i < jb.n_train; i += blockDim.x) {
            double s = 0, sm = 0;
            for (int c = 0; c < d; ++c) {
                double dl = yt[c] - static_cast<double>(tr[i * d + c]);
                s = fma(dl, dl, s);
                if (c == d - 2) sm = s;
            }
            aj += exp(-(s - mnj) * u2);
            if (ckde) am += exp(-(sm - mnm) * u2);
        }
        double tj = block_sum(aj, sh);
        double tm = 0;
        if (ckde) tm = block_sum(am, sh);
        if (threadIdx.x == 0) {
            double v = fj.lognorm_joint + log(tj) - mnj * u2;
            if (ckde) v -= fj.lognorm_marg + log(tm) - mnm * u2;
            out[fj.out_off + row] = v;
        }
    }
}

// sums[job] = sum of the job's logl entries (fixed order: deterministic)
__global__ void segsum_kernel(const PairJob* __restrict__ jobs, const FinJob* __restrict__ fin, const double* __restrict__ out,
                              double* __restrict__ sums) {
    __shared__ double sh[32];
    const long long m = jobs[blockIdx.x].m;
    const double* x = out + fin[blockIdx.x].out_off;
    double s = 0;
    for (long long i = threadIdx.x; i < m; i += blockDim.x) s += x[i];
    double tot = block_sum(s, sh);
    if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}
