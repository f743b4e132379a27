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
#else
#  ifndef TEGB_FN
#    error "teg_b200 kernels are CUDA-only: there is no host evaluation path"
#  endif
#endif

#endif /* TEG_B200_RUNTIME_CUH */
