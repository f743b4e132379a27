/*
 * teg_b200_runtime.cuh -- runtime declarations for generated CUDA (sm_100a) evaluation kernels.
 *
 * Plays the role of the reference's teg/include/teg_cuda_runtime.h:5-8 (TEG_DEVICE / TEG_HOST macros)
 * for the batched ``CUDA`` target.  The reference's ``generic_array<T,N>`` value type
 * (teg_generic_array.h:8-69) has no counterpart here: Tup-valued programs are scalarised by the
 * backend, so every component is an ordinary register and is stored to its own SoA output stream.
 *
 * Generated translation units are compiled with NVRTC for sm_100a with --fmad=false: the unit's
 * contract is "one IEEE operation per reference operation".  Do not add fast-math options.
 */
#ifndef TEG_B200_RUNTIME_CUH
#define TEG_B200_RUNTIME_CUH

#if defined(__CUDACC__)
#  ifndef TEGB_FN
#    define TEGB_FN static __device__ __forceinline__
#  endif
/* launch-invariant values, written by <name>_uniform and copied here by the runtime before <name>_main */
#  ifndef TEGB_UNIFORM_TABLE
#    define TEGB_UNIFORM_TABLE(T, name, n) __constant__ T name[n]
#  endif

/* Fixed-order block sum of K per-thread values (reverse-mode gradients w.r.t. shared parameters):
 * lanes are folded by a shuffle tree, warps by thread k in warp order; the result for component k goes to
 * partials[k * (number of blocks) + linear block index].  Warps whose lanes are all exactly zero (the
 * overwhelmingly common case: only instances crossed by a discontinuity have a non-zero gradient) skip
 * the shuffles -- adding zeros cannot change a sum, so the value is the same, bit for bit.  Deterministic:
 * depends only on the launch geometry and the values.  Must be reached by every thread of the block. */
template <typename T, int K, int BLOCK>
static __device__ __forceinline__ void tegb_block_sum_store(const T (&v)[K], T* __restrict__ partials,
                                                            const unsigned long long nblk,
                                                            const unsigned long long blk) {
    __shared__ T sm[K][BLOCK / 32];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    bool nz = false;
#pragma unroll
    for (int k = 0; k < K; k++) nz = nz || (v[k] != (T)0);
    const bool any = __any_sync(0xffffffffu, nz);
#pragma unroll
    for (int k = 0; k < K; k++) {
        T s = v[k];
        if (any) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s = s + __shfl_down_sync(0xffffffffu, s, off);
        }
        if (lane == 0) sm[k][warp] = s;
    }
    __syncthreads();
    if (threadIdx.x < K) {
        T t = sm[threadIdx.x][0];
#pragma unroll
        for (int w = 1; w < BLOCK / 32; w++) t = t + sm[threadIdx.x][w];
        partials[(unsigned long long)threadIdx.x * nblk + blk] = t;
    }
}

template <typename T, int K, int BLOCK>
static __device__ __forceinline__ void tegb_block_sum_store(const T (&v)[K], T* __restrict__ partials) {
    const unsigned long long nblk = (unsigned long long)gridDim.x * gridDim.y * gridDim.z;
    const unsigned long long blk = ((unsigned long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    tegb_block_sum_store<T, K, BLOCK>(v, partials, nblk, blk);
}

/* Sum over the TEAM threads that cooperate on one instance (TEAM consecutive threads, TEAM a power of two).
 * TEAM <= 32: xor butterfly inside the team's lanes.  Every stage adds the same two numbers on both partners,
 * so all members end with the same bits, and the combination order is fixed: deterministic.
 * TEAM  > 32: the team is the whole block; warp butterflies, then every thread adds the per-warp values in
 * warp order.  Reached by all members together: guards that enclose a team loop depend only on per-instance
 * values, which are identical across the team. */
template <typename T, int TEAM>
static __device__ __forceinline__ T tegb_team_sum(T v) {
    if (TEAM <= 32) {
        const unsigned lane = threadIdx.x & 31u;
        const unsigned mask = (TEAM == 32) ? 0xffffffffu : (((1u << (TEAM & 31)) - 1u) << (lane & ~(unsigned)(TEAM - 1)));
#pragma unroll
        for (int off = TEAM / 2; off > 0; off >>= 1) v = v + __shfl_xor_sync(mask, v, off);
        return v;
    } else {
        __shared__ T sm[32];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, off);
        __syncthreads();                       /* previous use of sm is over */
        if ((threadIdx.x & 31u) == 0) sm[threadIdx.x >> 5] = v;
        __syncthreads();
        T t = sm[0];
#pragma unroll
        for (int w = 1; w < TEAM / 32; w++) t = t + sm[w];
        return t;
    }
}
#else
#  ifndef TEGB_FN
#    error "teg_b200 kernels are CUDA-only: there is no host evaluation path"
#  endif
#endif

#endif /* TEG_B200_RUNTIME_CUH */
