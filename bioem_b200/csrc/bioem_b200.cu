// Context, batching and the device half of the C ABI (include/bioem_b200.h).
#include "../../include/bioem_b200.h"
#include "bioem_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace bioem;

#define BIOEM_SIZES(X) X(32) X(36) X(48) X(64) X(96) X(128) X(160) X(192) X(224) X(256) X(288) X(320) X(360) X(384) X(400)

static thread_local std::string g_err;
static int fail(int code, const std::string &msg)
{
  g_err = msg;
  return code;
}
#define CU(call)                                                                                          \
  do                                                                                                      \
  {                                                                                                       \
    cudaError_t e_ = (call);                                                                              \
    if (e_ != cudaSuccess)                                                                                \
      return fail(BIOEM_B200_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));               \
  } while (0)

struct bioem_b200_context
{
  bioem_b200_config cfg{};
  int device = 0;
  cudaStream_t stream = nullptr;
  int N = 0, nw = 0, nwp = 0, npos = 0;
  size_t map4 = 0; // float4 per packed map
  // inputs
  float4 *d_xyzr = nullptr;
  float *d_dens = nullptr;
  int A = 0;
  float NormDen = 0.f;
  float4 *d_angles = nullptr;
  int O = 0;
  float4 *d_ctf = nullptr;
  double *d_prior = nullptr;
  int C = 0;
  float4 *d_refs = nullptr;
  float *d_sumRef = nullptr, *d_sumsqRef = nullptr;
  int M = 0;
  float2 *d_tw_inv = nullptr, *d_tw_fwd = nullptr;
  unsigned char *d_wtab = nullptr;
  // batch buffers
  int OB = 0, OG = 1, nbands = 1, band_rows = 0;
  float *d_proj = nullptr;
  double *d_tempden = nullptr;
  float2 *d_scratch = nullptr;
  float4 *d_projfft = nullptr;
  float4 *d_conv = nullptr;
  ConvParam *d_cpar = nullptr;
  Running *d_partials = nullptr;
  size_t partials_cap = 0;
  // results
  Running *d_state = nullptr;
  ProbAngleOut *d_angtab = nullptr;
  ProbMapOut *d_out = nullptr;
  bool state_ready = false;
  // stats
  long long launches = 0, likelihoods = 0, lik_launches = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> lik_events;
  bool time_kernels = false;
};

// Device buffers come from the device's stream-ordered memory pool (cudaMallocAsync on the
// handle's stream): destroying a handle and creating the next one then costs microseconds per
// buffer instead of the ~20 ms of a cudaFree / cudaMalloc pair (measured: 0.3 s per handle).
static void dfree(bioem_b200_context *h, void *p)
{
  if (p)
    cudaFreeAsync(p, h->stream);
}

static void free_batch(bioem_b200_context *h)
{
  dfree(h, h->d_proj);
  dfree(h, h->d_tempden);
  dfree(h, h->d_scratch);
  dfree(h, h->d_projfft);
  dfree(h, h->d_conv);
  dfree(h, h->d_cpar);
  dfree(h, h->d_partials);
  h->d_proj = nullptr;
  h->d_tempden = nullptr;
  h->d_scratch = nullptr;
  h->d_projfft = nullptr;
  h->d_conv = nullptr;
  h->d_cpar = nullptr;
  h->d_partials = nullptr;
  h->OB = 0;
  h->partials_cap = 0;
}

template <int N> static size_t map4_of() { return Lay<N>::MAP4; }

static size_t map4_for(int N)
{
  switch (N)
  {
#define X(n)                                                                                              \
  case n:                                                                                                 \
    return map4_of<n>();
    BIOEM_SIZES(X)
#undef X
  }
  return 0;
}

template <int N> static void geo_of(int *r1, int *r2)
{
  *r1 = Lay<N>::R1;
  *r2 = Lay<N>::R2;
}
static void geo_for(int N, int *r1, int *r2)
{
  switch (N)
  {
#define X(n)                                                                                              \
  case n:                                                                                                 \
    geo_of<n>(r1, r2);                                                                                    \
    return;
    BIOEM_SIZES(X)
#undef X
  }
}

// ---------------------------------------------------------------------------------
// per-size launchers
// ---------------------------------------------------------------------------------
template <int N> static cudaError_t launch_pack(const float2 *src, float4 *dst, int nmaps, cudaStream_t s)
{
  dim3 grid(32, nmaps);
  pack_kernel<N><<<grid, 256, 0, s>>>(src, dst, nmaps);
  return cudaGetLastError();
}
template <int N> static cudaError_t launch_unpack(const float4 *src, float2 *dst, int nmaps, cudaStream_t s)
{
  dim3 grid(32, nmaps);
  unpack_kernel<N><<<grid, 256, 0, s>>>(src, dst, nmaps);
  return cudaGetLastError();
}
template <int N>
static cudaError_t launch_fft2d(const float *imgs, const double *tempden, int nbands, float normDen, const float2 *tw_fwd,
                                float2 *scratch, float4 *packed, int nimg, cudaStream_t s)
{
  using L = Lay<N>;
  dim3 g1((N / 2 + L::PC - 1) / L::PC, nimg);
  fft_rows_kernel<N><<<g1, NT, 0, s>>>(imgs, tempden, nbands, normDen, tw_fwd, scratch);
  dim3 g2((N / 2 + 1 + L::FKC - 1) / L::FKC, nimg);
  const size_t smem2 = sizeof(float2) * (size_t) (L::FKC * N + N);
  cudaError_t e = cudaFuncSetAttribute(fft_cols_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem2);
  if (e != cudaSuccess)
    return e;
  fft_cols_kernel<N><<<g2, NT, smem2, s>>>(scratch, tw_fwd, packed);
  return cudaGetLastError();
}
template <int N>
static cudaError_t launch_conv(const float4 *proj, const float4 *ctf, const double *prior, float4 *conv, ConvParam *cpar, int C,
                               int OBcur, float Nt, cudaStream_t s)
{
  dim3 g(C, OBcur);
  ctf_conv_kernel<N><<<g, NT, 0, s>>>(proj, ctf, prior, conv, cpar, C, Nt);
  return cudaGetLastError();
}
// dynamic + (an upper bound of the) static shared memory of the fused kernel
template <int N> static size_t lik_smem(int maxD, int nwp)
{
  return lik_smem_bytes<N>(maxD, nwp) + (size_t) lik_pending<N>() * LikSmem<N>::NWARP * 24 + 256;
}

static cudaError_t do_pack(int N, const float2 *src, float4 *dst, int nmaps, cudaStream_t s)
{
  switch (N)
  {
#define X(n)                                                                                              \
  case n:                                                                                                 \
    return launch_pack<n>(src, dst, nmaps, s);
    BIOEM_SIZES(X)
#undef X
  }
  return cudaErrorInvalidValue;
}
static cudaError_t do_unpack(int N, const float4 *src, float2 *dst, int nmaps, cudaStream_t s)
{
  switch (N)
  {
#define X(n)                                                                                              \
  case n:                                                                                                 \
    return launch_unpack<n>(src, dst, nmaps, s);
    BIOEM_SIZES(X)
#undef X
  }
  return cudaErrorInvalidValue;
}
static cudaError_t do_fft2d(int N, const float *imgs, const double *tempden, int nbands, float normDen, const float2 *tw,
                            float2 *scratch, float4 *packed, int nimg, cudaStream_t s)
{
  switch (N)
  {
#define X(n)                                                                                              \
  case n:                                                                                                 \
    return launch_fft2d<n>(imgs, tempden, nbands, normDen, tw, scratch, packed, nimg, s);
    BIOEM_SIZES(X)
#undef X
  }
  return cudaErrorInvalidValue;
}
static cudaError_t do_conv(int N, const float4 *proj, const float4 *ctf, const double *prior, float4 *conv, ConvParam *cpar,
                           int C, int OBcur, float Nt, cudaStream_t s)
{
  switch (N)
  {
#define X(n)                                                                                              \
  case n:                                                                                                 \
    return launch_conv<n>(proj, ctf, prior, conv, cpar, C, OBcur, Nt, s);
    BIOEM_SIZES(X)
#undef X
  }
  return cudaErrorInvalidValue;
}
// the fused kernel's variants live in one translation unit per image edge (lik_instance.inl)
#define X(n) extern "C" cudaError_t bioem_lik_launch_##n(const bioem::LikParams *, int, int, cudaStream_t);
BIOEM_SIZES(X)
#undef X
static cudaError_t do_lik(int N, const LikParams &p, int nblocks, int maxD, cudaStream_t s)
{
  switch (N)
  {
#define X(n)                                                                                              \
  case n:                                                                                                 \
    return bioem_lik_launch_##n(&p, nblocks, maxD, s);
    BIOEM_SIZES(X)
#undef X
  }
  return cudaErrorInvalidValue;
}
static size_t smem_for(int N, int maxD, int nwp)
{
  switch (N)
  {
#define X(n)                                                                                              \
  case n:                                                                                                 \
    return lik_smem<n>(maxD, nwp);
    BIOEM_SIZES(X)
#undef X
  }
  return 0;
}

// Gaussian priors on the CTF parameters (bioem_algorithm.h:49-67), per kernel, in double
static int set_ctf_priors(bioem_b200_context *h, const float *CtfParam4, int C)
{
  std::vector<double> prior(C);
  const bioem_b200_config &p = h->cfg;
  for (int c = 0; c < C; c++)
  {
    const float amp = CtfParam4[4 * c], pha = CtfParam4[4 * c + 1], env = CtfParam4[4 * c + 2];
    double pr;
    if (!p.tousepsf)
    {
      pr = env * env / 2. / p.sigmaPriorbctf / p.sigmaPriorbctf -
           (pha - p.Priordefcent) * (pha - p.Priordefcent) / 2. / p.sigmaPriordefo / p.sigmaPriordefo -
           (amp - p.Priorampcent) * (amp - p.Priorampcent) / 2. / p.sigmaPrioramp / p.sigmaPrioramp;
    }
    else
    {
      // PSF parameters are mapped to their Fourier-space counterparts first (bioem_algorithm.h:59-66)
      const double envF = 4. * M_PI * M_PI * env / (env * env + pha * pha);
      const double phaF = 4. * M_PI * M_PI * pha / (env * env + pha * pha);
      pr = envF * envF / 2. / p.sigmaPriorbctf / p.sigmaPriorbctf -
           (phaF - p.Priordefcent) * (phaF - p.Priordefcent) / 2. / p.sigmaPriordefo / p.sigmaPriordefo -
           (amp - p.Priorampcent) * (amp - p.Priorampcent) / 2. / p.sigmaPrioramp / p.sigmaPrioramp;
    }
    prior[c] = pr;
  }
  dfree(h, h->d_prior);
  h->d_prior = nullptr;
  CU(cudaMallocAsync((void **) &h->d_prior, sizeof(double) * C, h->stream));
  CU(cudaMemcpyAsync(h->d_prior, prior.data(), sizeof(double) * C, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  if (C != h->C)
    free_batch(h);
  h->C = C;
  return BIOEM_B200_OK;
}

// ---------------------------------------------------------------------------------
extern "C" {

const char *bioem_b200_last_error(void) { return g_err.c_str(); }
int bioem_b200_version(void) { return 100; }
int bioem_b200_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess)
  {
    cudaGetLastError();
    return 0;
  }
  return n;
}
int bioem_b200_supported_size(int N) { return map4_for(N) != 0; }

int bioem_b200_create(const bioem_b200_config *cfg, int device, bioem_b200_handle *out)
{
  if (!cfg || !out)
    return fail(BIOEM_B200_ERR_INVALID, "null argument");
  const int N = cfg->NumberPixels;
  if (!bioem_b200_supported_size(N))
    return fail(BIOEM_B200_ERR_INVALID, "NUMBER_PIXELS " + std::to_string(N) + " is not an instantiated image edge");
  if (cfg->GridSpaceCenter < 1 || cfg->maxDisplaceCenter < 0 || cfg->maxDisplaceCenter % cfg->GridSpaceCenter != 0)
    return fail(BIOEM_B200_ERR_INVALID, "DISPLACE_CENTER: grid spacing must be >= 1 and divide the maximum displacement");
  if (2 * cfg->maxDisplaceCenter + 1 > N)
    return fail(BIOEM_B200_ERR_INVALID, "DISPLACE_CENTER: window larger than the image");
  const int npos = cfg->maxDisplaceCenter / cfg->GridSpaceCenter + 1;
  const int nw = 2 * npos - 1;
  if (nw > 254)
    return fail(BIOEM_B200_ERR_INVALID, "displacement window has more than 254 points per axis");
  int ndev = bioem_b200_device_count();
  if (ndev == 0)
    return fail(BIOEM_B200_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
  if (device < 0 || device >= ndev)
    return fail(BIOEM_B200_ERR_INVALID, "bad device index");
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (smem_for(N, cfg->maxDisplaceCenter, nw + (nw & 1)) > (size_t) prop.sharedMemPerBlockOptin)
    return fail(BIOEM_B200_ERR_INVALID, "displacement window does not fit in shared memory for this image size");
  bioem_b200_context *h = new bioem_b200_context;
  h->cfg = *cfg;
  h->device = device;
  h->N = N;
  h->npos = npos;
  h->nw = nw;
  h->nwp = nw + (nw & 1);
  h->map4 = map4_for(N);
  h->time_kernels = getenv("BIOEM_B200_NO_KERNEL_TIMING") == nullptr;
  CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  {
    // keep freed blocks in the pool instead of returning them to the driver at every synchronisation
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess)
    {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
  }
  // twiddles (double -> float) at [n2*R1 + k1], and the displacement-window table
  int R1 = 0, R2 = 0;
  geo_for(N, &R1, &R2);
  std::vector<float2> twi(N), twf(N);
  for (int n2 = 0; n2 < R2; n2++)
    for (int k1 = 0; k1 < R1; k1++)
    {
      const long long pr = ((long long) n2 * k1) % N;
      const double ang = 2.0 * M_PI * (double) pr / (double) N;
      twi[n2 * R1 + k1] = make_float2((float) cos(ang), (float) sin(ang));
      twf[n2 * R1 + k1] = make_float2((float) cos(ang), (float) -sin(ang));
    }
  std::vector<unsigned char> wt(N, 255);
  for (int k = 0; k < npos; k++)
    wt[k * cfg->GridSpaceCenter] = (unsigned char) k;
  for (int k = 0; k < npos - 1; k++)
    wt[N - cfg->maxDisplaceCenter + k * cfg->GridSpaceCenter] = (unsigned char) (npos + k);
  CU(cudaMallocAsync((void **) &h->d_tw_inv, sizeof(float2) * N, h->stream));
  CU(cudaMallocAsync((void **) &h->d_tw_fwd, sizeof(float2) * N, h->stream));
  CU(cudaMallocAsync((void **) &h->d_wtab, N, h->stream));
  CU(cudaMemcpyAsync(h->d_tw_inv, twi.data(), sizeof(float2) * N, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  CU(cudaMemcpyAsync(h->d_tw_fwd, twf.data(), sizeof(float2) * N, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  CU(cudaMemcpyAsync(h->d_wtab, wt.data(), N, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  *out = h;
  return BIOEM_B200_OK;
}

int bioem_b200_destroy(bioem_b200_handle h)
{
  if (!h)
    return BIOEM_B200_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  free_batch(h);
  for (auto &ev : h->lik_events)
  {
    cudaEventDestroy(ev.first);
    cudaEventDestroy(ev.second);
  }
  dfree(h, h->d_xyzr);
  dfree(h, h->d_dens);
  dfree(h, h->d_angles);
  dfree(h, h->d_ctf);
  dfree(h, h->d_prior);
  dfree(h, h->d_refs);
  dfree(h, h->d_sumRef);
  dfree(h, h->d_sumsqRef);
  dfree(h, h->d_tw_inv);
  dfree(h, h->d_tw_fwd);
  dfree(h, h->d_wtab);
  dfree(h, h->d_state);
  dfree(h, h->d_angtab);
  dfree(h, h->d_out);
  cudaStreamSynchronize(h->stream);
  cudaStreamDestroy(h->stream);
  delete h;
  return BIOEM_B200_OK;
}

int bioem_b200_upload_model(bioem_b200_handle h, const bioem_b200_model_point *pts, int A, float NormDen)
{
  if (!h || !pts || A <= 0)
    return fail(BIOEM_B200_ERR_INVALID, "upload_model: bad argument");
  CU(cudaSetDevice(h->device));
  std::vector<float4> xyzr(A);
  std::vector<float> dens(A);
  for (int i = 0; i < A; i++)
  {
    xyzr[i] = make_float4(pts[i].pos[0], pts[i].pos[1], pts[i].pos[2], pts[i].radius);
    dens[i] = pts[i].density;
  }
  dfree(h, h->d_xyzr);
  dfree(h, h->d_dens);
  CU(cudaMallocAsync((void **) &h->d_xyzr, sizeof(float4) * A, h->stream));
  CU(cudaMallocAsync((void **) &h->d_dens, sizeof(float) * A, h->stream));
  CU(cudaMemcpyAsync(h->d_xyzr, xyzr.data(), sizeof(float4) * A, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  CU(cudaMemcpyAsync(h->d_dens, dens.data(), sizeof(float) * A, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  h->A = A;
  h->NormDen = NormDen;
  return BIOEM_B200_OK;
}

int bioem_b200_upload_orientations(bioem_b200_handle h, const float *angles4, int O)
{
  if (!h || !angles4 || O <= 0)
    return fail(BIOEM_B200_ERR_INVALID, "upload_orientations: bad argument");
  CU(cudaSetDevice(h->device));
  dfree(h, h->d_angles);
  CU(cudaMallocAsync((void **) &h->d_angles, sizeof(float4) * O, h->stream));
  CU(cudaMemcpyAsync(h->d_angles, angles4, sizeof(float4) * O, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  if (O != h->O)
  {
    dfree(h, h->d_angtab);
    h->d_angtab = nullptr;
    h->state_ready = false;
  }
  h->O = O;
  return BIOEM_B200_OK;
}

int bioem_b200_upload_ctf(bioem_b200_handle h, const float *refCTF, const float *CtfParam4, int C)
{
  if (!h || !refCTF || !CtfParam4 || C <= 0)
    return fail(BIOEM_B200_ERR_INVALID, "upload_ctf: bad argument");
  CU(cudaSetDevice(h->device));
  const int N = h->N;
  const size_t stdsz = (size_t) N * (N / 2 + 1);
  float2 *tmp = nullptr;
  CU(cudaMallocAsync((void **) &tmp, sizeof(float2) * stdsz * C, h->stream));
  CU(cudaMemcpyAsync(tmp, refCTF, sizeof(float2) * stdsz * C, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  dfree(h, h->d_ctf);
  h->d_ctf = nullptr;
  CU(cudaMallocAsync((void **) &h->d_ctf, sizeof(float4) * h->map4 * C, h->stream));
  CU(do_pack(N, tmp, h->d_ctf, C, h->stream));
  h->launches++;
  CU(cudaStreamSynchronize(h->stream));
  dfree(h, tmp);
  return set_ctf_priors(h, CtfParam4, C);
}

int bioem_b200_upload_ctf_real(bioem_b200_handle h, const float *kernels, const float *CtfParam4, int C)
{
  if (!h || !kernels || !CtfParam4 || C <= 0)
    return fail(BIOEM_B200_ERR_INVALID, "upload_ctf_real: bad argument");
  CU(cudaSetDevice(h->device));
  const int N = h->N;
  const size_t n2 = (size_t) N * N;
  float *d_img = nullptr;
  float2 *d_scr = nullptr;
  CU(cudaMallocAsync((void **) &d_img, sizeof(float) * n2 * C, h->stream));
  CU(cudaMallocAsync((void **) &d_scr, sizeof(float2) * (size_t) N * (N / 2 + 1) * C, h->stream));
  CU(cudaMemcpyAsync(d_img, kernels, sizeof(float) * n2 * C, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  dfree(h, h->d_ctf);
  h->d_ctf = nullptr;
  CU(cudaMallocAsync((void **) &h->d_ctf, sizeof(float4) * h->map4 * C, h->stream));
  CU(do_fft2d(N, d_img, nullptr, 0, 0.f, h->d_tw_fwd, d_scr, h->d_ctf, C, h->stream));
  h->launches += 2;
  CU(cudaStreamSynchronize(h->stream));
  dfree(h, d_img);
  dfree(h, d_scr);
  return set_ctf_priors(h, CtfParam4, C);
}

static int set_particle_count(bioem_b200_context *h, int M)
{
  if (M != h->M)
  {
    dfree(h, h->d_refs);
    dfree(h, h->d_sumRef);
    dfree(h, h->d_sumsqRef);
    dfree(h, h->d_state);
    dfree(h, h->d_out);
    dfree(h, h->d_angtab);
    h->d_refs = nullptr;
    h->d_sumRef = h->d_sumsqRef = nullptr;
    h->d_state = nullptr;
    h->d_out = nullptr;
    h->d_angtab = nullptr;
    h->state_ready = false;
    free_batch(h);
    CU(cudaMallocAsync((void **) &h->d_refs, sizeof(float4) * h->map4 * M, h->stream));
    CU(cudaMallocAsync((void **) &h->d_sumRef, sizeof(float) * M, h->stream));
    CU(cudaMallocAsync((void **) &h->d_sumsqRef, sizeof(float) * M, h->stream));
    CU(cudaMallocAsync((void **) &h->d_state, sizeof(Running) * M, h->stream));
    CU(cudaMallocAsync((void **) &h->d_out, sizeof(ProbMapOut) * M, h->stream));
    h->M = M;
  }
  return BIOEM_B200_OK;
}

int bioem_b200_upload_particles(bioem_b200_handle h, const float *maps, int M)
{
  if (!h || !maps || M <= 0)
    return fail(BIOEM_B200_ERR_INVALID, "upload_particles: bad argument");
  CU(cudaSetDevice(h->device));
  int rc = set_particle_count(h, M);
  if (rc)
    return rc;
  const int N = h->N;
  const size_t n2 = (size_t) N * N;
  // chunked so that the real-space staging stays small next to 180 GB of HBM
  const int chunk = (int) std::max<size_t>(1, std::min<size_t>((size_t) M, ((size_t) 1 << 30) / (n2 * 4)));
  float *d_img = nullptr;
  float2 *d_scr = nullptr;
  CU(cudaMallocAsync((void **) &d_img, sizeof(float) * n2 * chunk, h->stream));
  CU(cudaMallocAsync((void **) &d_scr, sizeof(float2) * (size_t) N * (N / 2 + 1) * chunk, h->stream));
  for (int m0 = 0; m0 < M; m0 += chunk)
  {
    const int mc = std::min(chunk, M - m0);
    CU(cudaMemcpyAsync(d_img, maps + (size_t) m0 * n2, sizeof(float) * n2 * mc, cudaMemcpyHostToDevice, h->stream));
    image_sums_kernel<<<(mc + 63) / 64, 64, 0, h->stream>>>(d_img, (int) n2, mc, h->d_sumRef + m0, h->d_sumsqRef + m0);
    CU(cudaGetLastError());
    CU(do_fft2d(N, d_img, nullptr, 0, 0.f, h->d_tw_fwd, d_scr, h->d_refs + (size_t) m0 * h->map4, mc, h->stream));
    h->launches += 3;
    CU(cudaStreamSynchronize(h->stream));
  }
  dfree(h, d_img);
  dfree(h, d_scr);
  return BIOEM_B200_OK;
}

int bioem_b200_upload_particles_mrc(bioem_b200_handle h, const float *raw, int M, int normalise)
{
  if (!h || !raw || M <= 0)
    return fail(BIOEM_B200_ERR_INVALID, "upload_particles_mrc: bad argument");
  CU(cudaSetDevice(h->device));
  int rc = set_particle_count(h, M);
  if (rc)
    return rc;
  const int N = h->N;
  const size_t n2 = (size_t) N * N;
  const int chunk = (int) std::max<size_t>(1, std::min<size_t>((size_t) M, ((size_t) 1 << 29) / (n2 * 4)));
  float *d_raw = nullptr, *d_img = nullptr, *d_mean = nullptr, *d_dev = nullptr;
  float2 *d_scr = nullptr;
  CU(cudaMallocAsync((void **) &d_raw, sizeof(float) * n2 * chunk, h->stream));
  CU(cudaMallocAsync((void **) &d_img, sizeof(float) * n2 * chunk, h->stream));
  CU(cudaMallocAsync((void **) &d_mean, sizeof(float) * chunk, h->stream));
  CU(cudaMallocAsync((void **) &d_dev, sizeof(float) * chunk, h->stream));
  CU(cudaMallocAsync((void **) &d_scr, sizeof(float2) * (size_t) N * (N / 2 + 1) * chunk, h->stream));
  for (int m0 = 0; m0 < M; m0 += chunk)
  {
    const int mc = std::min(chunk, M - m0);
    CU(cudaMemcpyAsync(d_raw, raw + (size_t) m0 * n2, sizeof(float) * n2 * mc, cudaMemcpyHostToDevice, h->stream));
    if (normalise)
      mrc_stats_kernel<<<(mc + 63) / 64, 64, 0, h->stream>>>(d_raw, (int) n2, mc, d_mean, d_dev);
    dim3 tg((N + 31) / 32, (N + 31) / 32, mc);
    mrc_transpose_kernel<<<tg, dim3(32, 8), 0, h->stream>>>(d_raw, d_mean, d_dev, N, normalise, d_img);
    image_sums_kernel<<<(mc + 63) / 64, 64, 0, h->stream>>>(d_img, (int) n2, mc, h->d_sumRef + m0, h->d_sumsqRef + m0);
    CU(cudaGetLastError());
    CU(do_fft2d(N, d_img, nullptr, 0, 0.f, h->d_tw_fwd, d_scr, h->d_refs + (size_t) m0 * h->map4, mc, h->stream));
    h->launches += 5;
    CU(cudaStreamSynchronize(h->stream));
  }
  dfree(h, d_raw);
  dfree(h, d_img);
  dfree(h, d_mean);
  dfree(h, d_dev);
  dfree(h, d_scr);
  return BIOEM_B200_OK;
}

int bioem_b200_upload_particles_fft(bioem_b200_handle h, const float *fft, const float *sum, const float *sumsq, int M)
{
  if (!h || !fft || !sum || !sumsq || M <= 0)
    return fail(BIOEM_B200_ERR_INVALID, "upload_particles_fft: bad argument");
  CU(cudaSetDevice(h->device));
  int rc = set_particle_count(h, M);
  if (rc)
    return rc;
  const int N = h->N;
  const size_t stdsz = (size_t) N * (N / 2 + 1);
  const int chunk = (int) std::max<size_t>(1, std::min<size_t>((size_t) M, ((size_t) 1 << 30) / (stdsz * 8)));
  float2 *tmp = nullptr;
  CU(cudaMallocAsync((void **) &tmp, sizeof(float2) * stdsz * chunk, h->stream));
  for (int m0 = 0; m0 < M; m0 += chunk)
  {
    const int mc = std::min(chunk, M - m0);
    CU(cudaMemcpyAsync(tmp, fft + (size_t) m0 * stdsz * 2, sizeof(float2) * stdsz * mc, cudaMemcpyHostToDevice, h->stream));
    CU(do_pack(N, tmp, h->d_refs + (size_t) m0 * h->map4, mc, h->stream));
    h->launches++;
    CU(cudaStreamSynchronize(h->stream));
  }
  dfree(h, tmp);
  CU(cudaMemcpyAsync(h->d_sumRef, sum, sizeof(float) * M, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  CU(cudaMemcpyAsync(h->d_sumsqRef, sumsq, sizeof(float) * M, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return BIOEM_B200_OK;
}

static int ensure_batch(bioem_b200_context *h)
{
  if (h->OB > 0)
    return BIOEM_B200_OK;
  const int N = h->N;
  const size_t mapbytes = h->map4 * sizeof(float4);
  // Orientations per batch: the conv spectra of a batch live in HBM (1 GB budget); what has to stay
  // in L2 is only the group of orientations the resident CTAs are working on (CTAs are numbered
  // image-fastest, so all images pass over one group before the next), i.e. a few MB.  Measured on
  // cfg2: 15 orientations per launch 18.93, 60: 19.21, 150: 19.26 M likelihoods/s (fewer launch tails).
  size_t budget = (size_t) 1 << 30;
  if (getenv("BIOEM_B200_CONV_MB"))
    budget = (size_t) atol(getenv("BIOEM_B200_CONV_MB")) << 20;
  long ob = (long) (budget / (mapbytes * (size_t) h->C));
  ob = std::max<long>(1, std::min<long>(ob, h->O));
  if (getenv("BIOEM_B200_OB"))
    ob = std::max<long>(1, std::min<long>(atol(getenv("BIOEM_B200_OB")), h->O));
  h->OB = (int) ob;
  // orientations per CTA: amortise the CTA prologue (and the one likelihood per CTA whose first radix pass
  // cannot be run ahead) over >= 64 likelihoods, keep >= 4 waves (cfg2: 1 -> 2 orientations, +0.5 %)
  // (above N = 224, one CTA per SM, more than one orientation per CTA at 32 CTFs costs 3-5 %)
  const int per_cta = h->N <= 224 ? 64 : 16;
  int og = std::max(1, (per_cta + h->C - 1) / h->C);
  const long long cta_slots = h->N <= 128 ? 592 : h->N <= 224 ? 296 : 148; // resident CTAs of the fused kernel per GPU
  while (og > 1 && (long long) h->M * ((h->OB + og - 1) / og) < 4LL * cta_slots)
    og--;
  if (getenv("BIOEM_B200_OG"))
    og = std::max(1, atoi(getenv("BIOEM_B200_OG")));
  h->OG = og;
  // bands of image rows per projection CTA: small bands = many CTAs (an orientation batch is only
  // ~15 images), at the price of every warp skipping more model points that miss its rows
  size_t band_budget = 24 * 1024;
  if (getenv("BIOEM_B200_BAND_KB"))
    band_budget = (size_t) atol(getenv("BIOEM_B200_BAND_KB")) * 1024;
  h->nbands = (int) (((size_t) N * N * 4 + band_budget - 1) / band_budget);
  h->band_rows = (N + h->nbands - 1) / h->nbands;
  h->nbands = (N + h->band_rows - 1) / h->band_rows;
  CU(cudaMallocAsync((void **) &h->d_proj, sizeof(float) * (size_t) N * N * h->OB, h->stream));
  CU(cudaMallocAsync((void **) &h->d_tempden, sizeof(double) * h->nbands * h->OB, h->stream));
  CU(cudaMallocAsync((void **) &h->d_scratch, sizeof(float2) * (size_t) N * (N / 2 + 1) * h->OB, h->stream));
  CU(cudaMallocAsync((void **) &h->d_projfft, mapbytes * h->OB, h->stream));
  CU(cudaMallocAsync((void **) &h->d_conv, mapbytes * (size_t) h->OB * h->C, h->stream));
  CU(cudaMallocAsync((void **) &h->d_cpar, sizeof(ConvParam) * (size_t) h->OB * h->C, h->stream));
  h->partials_cap = (size_t) h->M * ((h->OB + h->OG - 1) / h->OG);
  CU(cudaMallocAsync((void **) &h->d_partials, sizeof(Running) * h->partials_cap, h->stream));
  return BIOEM_B200_OK;
}

int bioem_b200_reset(bioem_b200_handle h)
{
  if (!h || h->M <= 0)
    return fail(BIOEM_B200_ERR_STATE, "reset: upload particles first");
  CU(cudaSetDevice(h->device));
  init_state_kernel<<<(h->M + 127) / 128, 128, 0, h->stream>>>(h->d_state, h->M);
  CU(cudaGetLastError());
  h->launches++;
  if (h->cfg.writeAngles)
  {
    if (h->O <= 0)
      return fail(BIOEM_B200_ERR_STATE, "reset: upload orientations first");
    const size_t n = (size_t) h->O * h->M;
    if (!h->d_angtab)
      CU(cudaMallocAsync((void **) &h->d_angtab, sizeof(ProbAngleOut) * n, h->stream));
    init_angles_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, h->stream>>>(h->d_angtab, n);
    CU(cudaGetLastError());
    h->launches++;
  }
  for (auto &ev : h->lik_events)
  {
    cudaEventDestroy(ev.first);
    cudaEventDestroy(ev.second);
  }
  h->lik_events.clear();
  h->launches = 0;
  h->likelihoods = 0;
  h->lik_launches = 0;
  h->state_ready = true;
  return BIOEM_B200_OK;
}

// stages 1 + 2 for orientations [o0, o0+OBcur) into the batch buffers
static int run_front(bioem_b200_context *h, int o0, int OBcur)
{
  const int N = h->N;
  ProjParams pp;
  pp.xyzr = h->d_xyzr;
  pp.dens = h->d_dens;
  pp.angles = h->d_angles;
  pp.proj = h->d_proj;
  pp.tempden = h->d_tempden;
  pp.skipped = nullptr;
  pp.A = h->A;
  pp.N = N;
  pp.band_rows = h->band_rows;
  pp.nbands = h->nbands;
  pp.o_base = o0;
  pp.doquater = h->cfg.doquater;
  pp.shiftX = h->cfg.shiftX;
  pp.shiftY = h->cfg.shiftY;
  pp.pixelSize = h->cfg.pixelSize;
  dim3 pg(OBcur, h->nbands);
  // the attribute belongs to the function, not to this handle: other handles (other image sizes)
  // may have changed it since, so it is set at every launch
  CU(cudaFuncSetAttribute(project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) ((size_t) h->band_rows * N * 4)));
  project_kernel<<<pg, 256, (size_t) h->band_rows * N * 4, h->stream>>>(pp);
  CU(cudaGetLastError());
  CU(do_fft2d(N, h->d_proj, h->d_tempden, h->nbands, h->NormDen, h->d_tw_fwd, h->d_scratch, h->d_projfft, OBcur, h->stream));
  CU(do_conv(N, h->d_projfft, h->d_ctf, h->d_prior, h->d_conv, h->d_cpar, h->C, OBcur, h->cfg.Ntotpi, h->stream));
  h->launches += 4;
  return BIOEM_B200_OK;
}

static void fill_lik_params(bioem_b200_context *h, LikParams &lp, int o0, int OBcur)
{
  const int N = h->N;
  lp.convs = h->d_conv;
  lp.refs = h->d_refs;
  lp.cpar = h->d_cpar;
  lp.sumRef = h->d_sumRef;
  lp.sumsqRef = h->d_sumsqRef;
  lp.tw_inv = h->d_tw_inv;
  lp.wtab = h->d_wtab;
  lp.partials = h->d_partials;
  lp.angles = h->cfg.writeAngles ? h->d_angtab : nullptr;
  lp.dbg_values = nullptr;
  lp.M = h->M;
  lp.C = h->C;
  lp.OBcur = OBcur;
  lp.OG = h->OG;
  lp.o_base = o0;
  lp.nw = h->nw;
  lp.nwp = h->nwp;
  lp.Ntotpi = h->cfg.Ntotpi;
  lp.invNN = 1.0f / (float) (N * N);
  lp.acoef_d = (double) (3.f - h->cfg.Ntotpi) * 0.5;
  lp.ex2coef = (float) (lp.acoef_d * 1.4426950408889634074);
}

int bioem_b200_run(bioem_b200_handle h, int oBegin, int oEnd)
{
  if (!h)
    return fail(BIOEM_B200_ERR_INVALID, "run: null handle");
  if (h->A <= 0 || h->O <= 0 || h->C <= 0 || h->M <= 0)
    return fail(BIOEM_B200_ERR_STATE, "run: model, orientations, CTF table and particles must be uploaded first");
  if (oBegin < 0 || oEnd > h->O || oBegin > oEnd)
    return fail(BIOEM_B200_ERR_INVALID, "run: orientation range out of bounds");
  CU(cudaSetDevice(h->device));
  if (!h->state_ready)
  {
    int rc = bioem_b200_reset(h);
    if (rc)
      return rc;
  }
  int rc = ensure_batch(h);
  if (rc)
    return rc;
  for (int o0 = oBegin; o0 < oEnd; o0 += h->OB)
  {
    const int OBcur = std::min(h->OB, oEnd - o0);
    rc = run_front(h, o0, OBcur);
    if (rc)
      return rc;
    LikParams lp;
    fill_lik_params(h, lp, o0, OBcur);
    const int NG = (OBcur + h->OG - 1) / h->OG;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (h->time_kernels)
    {
      CU(cudaEventCreate(&e0));
      CU(cudaEventCreate(&e1));
      CU(cudaEventRecord(e0, h->stream));
    }
    CU(do_lik(h->N, lp, h->M * NG, h->cfg.maxDisplaceCenter, h->stream));
    if (h->time_kernels)
    {
      CU(cudaEventRecord(e1, h->stream));
      h->lik_events.emplace_back(e0, e1);
    }
    merge_partials_kernel<<<(h->M + 127) / 128, 128, 0, h->stream>>>(h->d_partials, NG, h->M, h->d_state);
    CU(cudaGetLastError());
    h->launches += 2;
    h->lik_launches += 1;
    h->likelihoods += (long long) OBcur * h->C * h->M;
  }
  return BIOEM_B200_OK;
}

int bioem_b200_synchronize(bioem_b200_handle h)
{
  if (!h)
    return fail(BIOEM_B200_ERR_INVALID, "null handle");
  CU(cudaSetDevice(h->device));
  CU(cudaStreamSynchronize(h->stream));
  return BIOEM_B200_OK;
}

int bioem_b200_download(bioem_b200_handle h, bioem_b200_prob_map *maps_out, bioem_b200_prob_angle *angles_out)
{
  if (!h || !maps_out)
    return fail(BIOEM_B200_ERR_INVALID, "download: bad argument");
  if (!h->state_ready)
    return fail(BIOEM_B200_ERR_STATE, "download: nothing has been run");
  CU(cudaSetDevice(h->device));
  static_assert(sizeof(ProbMapOut) == sizeof(bioem_b200_prob_map) && sizeof(ProbMapOut) == 40, "result layout");
  static_assert(sizeof(ProbAngleOut) == sizeof(bioem_b200_prob_angle) && sizeof(ProbAngleOut) == 16, "result layout");
  finalize_kernel<<<(h->M + 127) / 128, 128, 0, h->stream>>>(h->d_state, h->d_sumRef, h->M, h->nw, h->npos,
                                                            h->cfg.maxDisplaceCenter, h->cfg.GridSpaceCenter, h->cfg.Ntotpi,
                                                            h->d_out);
  CU(cudaGetLastError());
  h->launches++;
  CU(cudaMemcpyAsync(maps_out, h->d_out, sizeof(ProbMapOut) * h->M, cudaMemcpyDeviceToHost, h->stream));
  if (angles_out && h->cfg.writeAngles && h->d_angtab)
    CU(cudaMemcpyAsync(angles_out, h->d_angtab, sizeof(ProbAngleOut) * (size_t) h->O * h->M, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return BIOEM_B200_OK;
}

int bioem_b200_download_top_angles(bioem_b200_handle h, int oBegin, int oEnd, int K, bioem_b200_top_angle *out)
{
  if (!h || !out || K <= 0 || oBegin < 0 || oEnd > (h ? h->O : 0) || oBegin > oEnd)
    return fail(BIOEM_B200_ERR_INVALID, "download_top_angles: bad argument");
  if (!h->cfg.writeAngles || !h->d_angtab)
    return fail(BIOEM_B200_ERR_STATE, "download_top_angles: the handle was created with writeAngles == 0");
  if (!h->state_ready)
    return fail(BIOEM_B200_ERR_STATE, "download_top_angles: nothing has been run");
  static_assert(sizeof(TopAngleOut) == sizeof(bioem_b200_top_angle) && sizeof(TopAngleOut) == 24, "result layout");
  CU(cudaSetDevice(h->device));
  const size_t n = (size_t) h->M * K;
  double *d_key = nullptr;
  TopAngleOut *d_top = nullptr;
  CU(cudaMallocAsync((void **) &d_key, n * sizeof(double), h->stream));
  CU(cudaMallocAsync((void **) &d_top, n * sizeof(TopAngleOut), h->stream));
  top_angles_kernel<<<(h->M + 63) / 64, 64, 0, h->stream>>>(h->d_angtab, h->M, oBegin, oEnd, K, d_key, d_top);
  CU(cudaGetLastError());
  h->launches++;
  CU(cudaMemcpyAsync(out, d_top, n * sizeof(TopAngleOut), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaFreeAsync(d_key, h->stream));
  CU(cudaFreeAsync(d_top, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return BIOEM_B200_OK;
}

size_t bioem_b200_partial_bytes(bioem_b200_handle h) { return h ? sizeof(Running) * (size_t) h->M : 0; }

int bioem_b200_export_partial(bioem_b200_handle h, void *device_dst)
{
  if (!h || !device_dst || !h->state_ready)
    return fail(BIOEM_B200_ERR_STATE, "export_partial: nothing to export");
  CU(cudaSetDevice(h->device));
  CU(cudaMemcpyAsync(device_dst, h->d_state, sizeof(Running) * h->M, cudaMemcpyDeviceToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return BIOEM_B200_OK;
}

int bioem_b200_import_partials(bioem_b200_handle h, const void *device_gathered, int nRanks)
{
  if (!h || !device_gathered || nRanks <= 0)
    return fail(BIOEM_B200_ERR_INVALID, "import_partials: bad argument");
  CU(cudaSetDevice(h->device));
  init_state_kernel<<<(h->M + 127) / 128, 128, 0, h->stream>>>(h->d_state, h->M);
  merge_partials_kernel<<<(h->M + 127) / 128, 128, 0, h->stream>>>((const Running *) device_gathered, nRanks, h->M, h->d_state);
  CU(cudaGetLastError());
  h->launches += 2;
  h->state_ready = true;
  CU(cudaStreamSynchronize(h->stream));
  return BIOEM_B200_OK;
}

void *bioem_b200_stream(bioem_b200_handle h) { return h ? (void *) h->stream : nullptr; }
void *bioem_b200_device_angles(bioem_b200_handle h) { return h ? (void *) h->d_angtab : nullptr; }

int bioem_b200_stats(bioem_b200_handle h, long long *launches, long long *likelihoods)
{
  if (!h)
    return fail(BIOEM_B200_ERR_INVALID, "null handle");
  if (launches)
    *launches = h->launches;
  if (likelihoods)
    *likelihoods = h->likelihoods;
  return BIOEM_B200_OK;
}

int bioem_b200_kernel_time(bioem_b200_handle h, double *ms, long long *n)
{
  if (!h)
    return fail(BIOEM_B200_ERR_INVALID, "null handle");
  CU(cudaSetDevice(h->device));
  CU(cudaStreamSynchronize(h->stream));
  double t = 0.0;
  for (auto &ev : h->lik_events)
  {
    float f = 0.f;
    CU(cudaEventElapsedTime(&f, ev.first, ev.second));
    t += f;
  }
  if (ms)
    *ms = t;
  if (n)
    *n = (long long) h->lik_events.size();
  return BIOEM_B200_OK;
}

// ------------------------------------------------------------------ inspection
int bioem_b200_debug_projection(bioem_b200_handle h, int o, float *out)
{
  if (!h || !out || o < 0 || o >= h->O || h->A <= 0 || h->C <= 0 || h->M <= 0)
    return fail(BIOEM_B200_ERR_INVALID, "debug_projection: bad argument / inputs missing");
  CU(cudaSetDevice(h->device));
  int rc = ensure_batch(h);
  if (rc)
    return rc;
  rc = run_front(h, o, 1);
  if (rc)
    return rc;
  const size_t n2 = (size_t) h->N * h->N;
  std::vector<double> td(h->nbands);
  CU(cudaMemcpyAsync(out, h->d_proj, sizeof(float) * n2, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(td.data(), h->d_tempden, sizeof(double) * h->nbands, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  double t = 0.0;
  for (double v : td)
    t += v;
  const float ratio = h->NormDen / (float) t;
  for (size_t i = 0; i < n2; i++)
    out[i] *= ratio;
  return BIOEM_B200_OK;
}

int bioem_b200_debug_convolved(bioem_b200_handle h, int o, int c, float *conv_out, float *sumC, float *sumsqC)
{
  if (!h || o < 0 || o >= h->O || c < 0 || c >= h->C || h->A <= 0 || h->M <= 0)
    return fail(BIOEM_B200_ERR_INVALID, "debug_convolved: bad argument / inputs missing");
  CU(cudaSetDevice(h->device));
  int rc = ensure_batch(h);
  if (rc)
    return rc;
  rc = run_front(h, o, 1);
  if (rc)
    return rc;
  const size_t stdsz = (size_t) h->N * (h->N / 2 + 1);
  if (conv_out)
  {
    float2 *tmp = nullptr;
    CU(cudaMallocAsync((void **) &tmp, sizeof(float2) * stdsz, h->stream));
    CU(do_unpack(h->N, h->d_conv + (size_t) c * h->map4, tmp, 1, h->stream));
    CU(cudaMemcpyAsync(conv_out, tmp, sizeof(float2) * stdsz, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    dfree(h, tmp);
  }
  ConvParam cp;
  CU(cudaMemcpyAsync(&cp, h->d_cpar + c, sizeof(cp), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  if (sumC)
    *sumC = cp.sumC;
  if (sumsqC)
    *sumsqC = cp.sumsqC;
  return BIOEM_B200_OK;
}

int bioem_b200_debug_correlation(bioem_b200_handle h, int o, int c, int m, float *values, int *nvalues)
{
  if (!h || !values || o < 0 || o >= h->O || c < 0 || c >= h->C || m < 0 || m >= h->M || h->A <= 0)
    return fail(BIOEM_B200_ERR_INVALID, "debug_correlation: bad argument / inputs missing");
  CU(cudaSetDevice(h->device));
  int rc = ensure_batch(h);
  if (rc)
    return rc;
  rc = run_front(h, o, 1);
  if (rc)
    return rc;
  const size_t nv = (size_t) h->nw * h->nw;
  float *d_dbg = nullptr;
  Running *d_part = nullptr;
  CU(cudaMallocAsync((void **) &d_dbg, sizeof(float) * nv * h->C, h->stream));
  CU(cudaMallocAsync((void **) &d_part, sizeof(Running), h->stream));
  LikParams lp;
  fill_lik_params(h, lp, o, 1);
  lp.refs = h->d_refs + (size_t) m * h->map4;
  lp.sumRef = h->d_sumRef + m;
  lp.sumsqRef = h->d_sumsqRef + m;
  lp.partials = d_part;
  lp.angles = nullptr;
  lp.dbg_values = d_dbg;
  lp.M = 1;
  lp.OG = 1;
  CU(do_lik(h->N, lp, 1, h->cfg.maxDisplaceCenter, h->stream));
  CU(cudaMemcpyAsync(values, d_dbg + (size_t) c * nv, sizeof(float) * nv, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  dfree(h, d_dbg);
  dfree(h, d_part);
  if (nvalues)
    *nvalues = (int) nv;
  return BIOEM_B200_OK;
}

int bioem_b200_debug_particle(bioem_b200_handle h, int m, float *fft_out, float *sum, float *sumsq)
{
  if (!h || m < 0 || m >= h->M)
    return fail(BIOEM_B200_ERR_INVALID, "debug_particle: bad argument");
  CU(cudaSetDevice(h->device));
  const size_t stdsz = (size_t) h->N * (h->N / 2 + 1);
  if (fft_out)
  {
    float2 *tmp = nullptr;
    CU(cudaMallocAsync((void **) &tmp, sizeof(float2) * stdsz, h->stream));
    CU(do_unpack(h->N, h->d_refs + (size_t) m * h->map4, tmp, 1, h->stream));
    CU(cudaMemcpyAsync(fft_out, tmp, sizeof(float2) * stdsz, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    dfree(h, tmp);
  }
  if (sum)
    CU(cudaMemcpyAsync(sum, h->d_sumRef + m, sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  if (sumsq)
    CU(cudaMemcpyAsync(sumsq, h->d_sumsqRef + m, sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return BIOEM_B200_OK;
}

// host-side merge of per-rank results (same rule as merge_partials_kernel: strict '<'
// in rank order keeps the lowest rank on ties)
int bioem_b200_merge_host(const bioem_b200_prob_map *parts, int nRanks, int nMaps, bioem_b200_prob_map *out)
{
  if (!parts || !out || nRanks <= 0 || nMaps <= 0)
    return fail(BIOEM_B200_ERR_INVALID, "merge_host: bad argument");
  for (int m = 0; m < nMaps; m++)
  {
    bioem_b200_prob_map s = parts[m];
    for (int r = 1; r < nRanks; r++)
    {
      const bioem_b200_prob_map &p = parts[(size_t) r * nMaps + m];
      if (s.Constoadd < p.Constoadd)
      {
        const double T = s.Total * exp(s.Constoadd - p.Constoadd) + p.Total;
        s = p;
        s.Total = T;
      }
      else
        s.Total += p.Total * exp(p.Constoadd - s.Constoadd);
    }
    out[m] = s;
  }
  return BIOEM_B200_OK;
}

} // extern "C"
