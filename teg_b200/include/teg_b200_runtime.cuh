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

/* Fixed-order sum of K per-thread values over the 32 lanes of a warp (shuffle tree); lane 0 stores component k at
 * dst[k * stride].  Tiled render programs keep one partial sum per tile (a warp's 32 columns x the tile's rows): the
 * tile, not the block that happened to compute it, identifies the slot.  Warps whose lanes are all exactly zero skip
 * the shuffles (adding zeros cannot change a sum).  Reached by all 32 lanes. */
template <typename T, int K>
static __device__ __forceinline__ void tegb_warp_sum_store(const T (&v)[K], T* __restrict__ dst, const unsigned long long stride) {
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
        if ((threadIdx.x & 31u) == 0) dst[(unsigned long long)k * stride] = s;
    }
}

/* Teams: TEAM consecutive threads (a power of two) cooperate on one instance by splitting its OUTERMOST sample loop.
 * The reference folds the outer terms left to right (rectangular_rule.template.c:4-9), and so does the team: the
 * loop advances in chunks of TEAM * U consecutive samples, member r evaluates the U samples chunk + r * U + u (its
 * inner integrals are ordinary sequential folds), and then EVERY member adds the chunk's terms to its copy of the
 * accumulator in sample order.  All members hold the same bits at all times and the result equals the sequential
 * fold bit for bit -- no partial sums, nothing re-associated.
 * TEAM <= 32: the terms travel by shuffles inside the team's lanes.  TEAM > 32: the team is the whole block; the terms
 * go through shared memory.  `count` = number of valid samples in this chunk (a multiple of U; team-uniform).
 * Reached by all members together: guards that enclose a team loop depend only on per-instance values. */
template <typename T, int TEAM, int U>
static __device__ __forceinline__ T tegb_team_fold(T acc, const T (&t)[U], const unsigned int count) {
    if (TEAM <= 32) {
        const unsigned lane = threadIdx.x & 31u;
        const unsigned first = lane & ~(unsigned)(TEAM - 1);
        const unsigned mask = (TEAM == 32) ? 0xffffffffu : (((1u << (TEAM & 31)) - 1u) << first);
        for (unsigned int l = 0; l * U < count; ++l) {
#pragma unroll
            for (int u = 0; u < U; u++) acc = acc + __shfl_sync(mask, t[u], (int)(first + l));
        }
        return acc;
    } else {
        __shared__ T sm[TEAM * U];
        __syncthreads();                       /* previous use of sm is over */
#pragma unroll
        for (int u = 0; u < U; u++) sm[threadIdx.x * U + u] = t[u];
        __syncthreads();
        for (unsigned int i = 0; i < count; ++i) acc = acc + sm[i];
        return acc;
    }
}
#else
#  ifndef TEGB_FN
#    error "teg_b200 kernels are CUDA-only: there is no host evaluation path"
#  endif
#endif

#endif /* TEG_B200_RUNTIME_CUH */
