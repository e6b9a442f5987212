// One translation unit per image edge (BIOEM_N), so the fused-kernel variants compile in
// parallel.  Exports  bioem_lik_launch_<N>(params, nblocks, maxDisplaceCenter, stream).
#define BIOEM_LIK_ONLY
#include "bioem_kernels.cuh"

namespace bioem
{
template <int W, bool ZM> static cudaError_t lik_launch_wz(const LikParams &p, int nblocks, cudaStream_t s)
{
  const size_t smem = LikSmem<BIOEM_N>::bytes(W, p.nwp);
  // attribute is per device context: set on every launch (cheap next to the kernel)
  cudaError_t e = cudaFuncSetAttribute(likelihood_kernel<BIOEM_N, W, ZM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  if (e != cudaSuccess)
    return e;
  likelihood_kernel<BIOEM_N, W, ZM><<<nblocks, LikSmem<BIOEM_N>::LNT, smem, s>>>(p);
  return cudaGetLastError();
}
// p.zmode selects the cached-product variant (real CTF tables, LikParams::projs / kreal / zbuf)
template <int W> static cudaError_t lik_launch_w(const LikParams &p, int nblocks, cudaStream_t s)
{
  return p.zmode ? lik_launch_wz<W, true>(p, nblocks, s) : lik_launch_wz<W, false>(p, nblocks, s);
}
} // namespace bioem

#define BIOEM_CAT2(a, b) a##b
#define BIOEM_CAT(a, b) BIOEM_CAT2(a, b)

extern "C" cudaError_t BIOEM_CAT(bioem_lik_launch_, BIOEM_N)(const bioem::LikParams *p, int nblocks, int maxD, cudaStream_t s)
{
  using L = bioem::Lay<BIOEM_N>;
  constexpr int HALF = L::R2 / 2;
  // radix-R2 output groups that can hold a displacement in [-maxD, maxD] (lik_window_groups)
  const int w = bioem::lik_window_groups<BIOEM_N>(maxD);
#ifdef BIOEM_ONLY_W // experiments (tools/build_variant.py): compile a single window-group variant
  return w == BIOEM_ONLY_W ? bioem::lik_launch_w<BIOEM_ONLY_W>(*p, nblocks, s) : cudaErrorInvalidValue;
#else
  if constexpr (L::G::GENERIC)
    return bioem::lik_launch_w<(HALF > 0 ? HALF : 1)>(*p, nblocks, s);
  else
  {
  if (w == 1 && HALF > 1)
    return bioem::lik_launch_w<1>(*p, nblocks, s);
  if constexpr (HALF > 2)
    if (w == 2)
      return bioem::lik_launch_w<2>(*p, nblocks, s);
  if constexpr (HALF > 3)
    if (w == 3)
      return bioem::lik_launch_w<3>(*p, nblocks, s);
  if constexpr (HALF > 4)
    if (w == 4)
      return bioem::lik_launch_w<4>(*p, nblocks, s);
  if constexpr (HALF > 5)
    if (w == 5)
      return bioem::lik_launch_w<5>(*p, nblocks, s);
  if constexpr (HALF > 6)
    if (w == 6)
      return bioem::lik_launch_w<6>(*p, nblocks, s);
  if constexpr (HALF > 7)
    if (w == 7)
      return bioem::lik_launch_w<7>(*p, nblocks, s);
  return bioem::lik_launch_w<HALF>(*p, nblocks, s);
  }
#endif
}
