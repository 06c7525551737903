// HBM-bound movers of the forward path (NHWC bf16, 16-byte vector accesses, coalesced):
// Focus space-to-depth gather, SPP max-pools, nearest 2x upsample, Add/Add2, Concat copy,
// plus the plain CUDA-core conv used only to cross-check the tcgen05 kernel in tests.
#include "cft_common.cuh"

namespace {
using namespace cft;

constexpr int kThreads = 256;

inline int grid_for(long long work, int threads, int max_blocks_per_sm = 16) {
  long long blocks = (work + threads - 1) / threads;
  const long long cap = static_cast<long long>(sm_count()) * max_blocks_per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

// ------------------------------------------------------------------ Focus gather
// One thread per output pixel: reads a 2x2 patch of each of the 3 planes (two float2 / bf162
// loads per plane, consecutive threads -> consecutive addresses), writes 32 B (16 bf16).
template <typename T>
__device__ __forceinline__ void focus_patch(const T* __restrict__ img, long long bstride, int b, int H, int W, int oy, int ox,
                                            float* f /*[12]*/) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const T* p0 = img + static_cast<long long>(b) * bstride + (static_cast<long long>(c) * H + 2 * oy) * W + 2 * ox;
    const T* p1 = p0 + W;
    float a00, a01, a10, a11;
    if constexpr (sizeof(T) == 1) {
      const uchar2 r0 = *reinterpret_cast<const uchar2*>(p0);
      const uchar2 r1 = *reinterpret_cast<const uchar2*>(p1);
      a00 = r0.x / 255.0f; a01 = r0.y / 255.0f; a10 = r1.x / 255.0f; a11 = r1.y / 255.0f;
    } else if constexpr (sizeof(T) == 4) {
      const float2 r0 = *reinterpret_cast<const float2*>(p0);
      const float2 r1 = *reinterpret_cast<const float2*>(p1);
      a00 = r0.x; a01 = r0.y; a10 = r1.x; a11 = r1.y;
    } else {
      const float2 r0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p0));
      const float2 r1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p1));
      a00 = r0.x; a01 = r0.y; a10 = r1.x; a11 = r1.y;
    }
    // channel = (dy + 2*dx)*3 + c   (models/common.py:179: [::2,::2],[1::2,::2],[::2,1::2],[1::2,1::2])
    f[0 * 3 + c] = a00;
    f[1 * 3 + c] = a10;
    f[2 * 3 + c] = a01;
    f[3 * 3 + c] = a11;
  }
}

// kLayout 0: one thread per output pixel, 16 channels (32 B).
// kLayout 1: four threads per output pixel, each writes 32 B of the 128-byte pixel: slots 0..2 = the patch of
// x-1, x, x+1 (x-direction im2col), slot 3 = zeros.  A warp writes 8 pixels = 1 KiB contiguous.
template <typename T, int kLayout>
__global__ void focus_gather_kernel(const T* __restrict__ img, __nv_bfloat16* __restrict__ y, int B, int H, int W,
                                    long long bstride) {
  pdl_prologue();
  const int Ho = H / 2, Wo = W / 2;
  const long long npix = static_cast<long long>(B) * Ho * Wo;
  const long long total = kLayout == 0 ? npix : npix * 4;
  for (long long it = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; it < total;
       it += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long idx = kLayout == 0 ? it : (it >> 2);
    const int slot = kLayout == 0 ? 0 : static_cast<int>(it & 3);
    const int ox = static_cast<int>(idx % Wo);
    long long t = idx / Wo;
    const int oy = static_cast<int>(t % Ho);
    const int b = static_cast<int>(t / Ho);
    float f[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = 0.f;
    if constexpr (kLayout == 0) {
      focus_patch(img, bstride, b, H, W, oy, ox, f);
      bf16x8* out = reinterpret_cast<bf16x8*>(y + idx * 16);
      out[0] = pack8(f);
      out[1] = pack8(f + 8);
    } else {
      const int xx = ox + slot - 1;
      if (slot < 3 && xx >= 0 && xx < Wo) focus_patch(img, bstride, b, H, W, oy, xx, f);
      bf16x8* out = reinterpret_cast<bf16x8*>(y + idx * 64 + slot * 16);
      out[0] = pack8(f);
      out[1] = pack8(f + 8);
    }
  }
}

// ------------------------------------------------------------------ max pool k x k, stride 1
__global__ void maxpool_s1_kernel(const __nv_bfloat16* __restrict__ x, int ldx, __nv_bfloat16* __restrict__ y, int ldy,
                                  int B, int H, int W, int C8, int k) {
  pdl_prologue();
  const long long total = static_cast<long long>(B) * H * W * C8;
  const int r = k / 2;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(idx % C8);
    long long t = idx / C8;
    const int ox = static_cast<int>(t % W);
    t /= W;
    const int oy = static_cast<int>(t % H);
    const int b = static_cast<int>(t / H);
    float m[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = -INFINITY;
    const int y0 = max(oy - r, 0), y1 = min(oy + r, H - 1);
    const int x0 = max(ox - r, 0), x1 = min(ox + r, W - 1);
    for (int yy = y0; yy <= y1; ++yy)
      for (int xx = x0; xx <= x1; ++xx) {
        float f[8];
        unpack8(*reinterpret_cast<const bf16x8*>(x + ((static_cast<long long>(b) * H + yy) * W + xx) * ldx + cv * 8), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], f[i]);
      }
    *reinterpret_cast<bf16x8*>(y + ((static_cast<long long>(b) * H + oy) * W + ox) * ldy + cv * 8) = pack8(m);
  }
}


// ------------------------------------------------------------------ SPP: cascade of stride-1 max pools, one pass
// One CTA per (image, 16-channel group): the HxW plane of 16 channels (32 B per pixel = one full sector) is
// staged in smem, then three separable k x k max pools are applied back to back (row pass + column pass each);
// the result of every stage is written to its own channel slice.  pool_9 = pool_5(pool_5), pool_13 = pool_5(pool_9)
// for stride-1 pools with -inf padding, so SPP's (5, 9, 13) is the cascade (5, 5, 5).
struct __align__(16) bf16x16 {
  bf16x8 lo, hi;
};
__device__ __forceinline__ bf16x8 max8(const bf16x8& a, const bf16x8& b) {
  bf16x8 r;
#pragma unroll
  for (int i = 0; i < 4; ++i) r.set(i, __hmax2(a.get(i), b.get(i)));
  return r;
}
__global__ void maxpool_cascade_kernel(const __nv_bfloat16* __restrict__ x, int ldx, __nv_bfloat16* __restrict__ y0,
                                       __nv_bfloat16* __restrict__ y1, __nv_bfloat16* __restrict__ y2, int ldy, int H,
                                       int W, int k0, int k1, int k2) {
  pdl_prologue();
  extern __shared__ uint8_t sm_raw[];
  bf16x16* bufA = reinterpret_cast<bf16x16*>(sm_raw);
  bf16x16* bufB = bufA + H * W;
  const int cg = blockIdx.x, b = blockIdx.y;
  const int npix = H * W;
  const long long img = static_cast<long long>(b) * npix;
  for (int i = threadIdx.x; i < npix; i += blockDim.x) {
    const __nv_bfloat16* p = x + (img + i) * ldx + cg * 16;
    bufA[i].lo = *reinterpret_cast<const bf16x8*>(p);
    bufA[i].hi = *reinterpret_cast<const bf16x8*>(p + 8);
  }
  __syncthreads();
  __nv_bfloat16* outs[3] = {y0, y1, y2};
  const int ks[3] = {k0, k1, k2};
  for (int st = 0; st < 3; ++st) {
    const int r = ks[st] / 2;
    for (int i = threadIdx.x; i < npix; i += blockDim.x) {      // row pass: A -> B
      const int yy = i / W, xx = i - yy * W;
      const int xa = max(xx - r, 0), xb = min(xx + r, W - 1);
      bf16x16 m = bufA[yy * W + xa];
      for (int q = xa + 1; q <= xb; ++q) {
        const bf16x16 v = bufA[yy * W + q];
        m.lo = max8(m.lo, v.lo);
        m.hi = max8(m.hi, v.hi);
      }
      bufB[i] = m;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < npix; i += blockDim.x) {      // column pass: B -> A (+ global)
      const int yy = i / W, xx = i - yy * W;
      const int ya = max(yy - r, 0), yb = min(yy + r, H - 1);
      bf16x16 m = bufB[ya * W + xx];
      for (int q = ya + 1; q <= yb; ++q) {
        const bf16x16 v = bufB[q * W + xx];
        m.lo = max8(m.lo, v.lo);
        m.hi = max8(m.hi, v.hi);
      }
      bufA[i] = m;
      __nv_bfloat16* o = outs[st] + (img + i) * ldy + cg * 16;
      *reinterpret_cast<bf16x8*>(o) = m.lo;
      *reinterpret_cast<bf16x8*>(o + 8) = m.hi;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ nearest 2x upsample
__global__ void upsample2x_kernel(const __nv_bfloat16* __restrict__ x, int ldx, __nv_bfloat16* __restrict__ y, int ldy,
                                  int B, int H, int W, int C8) {
  pdl_prologue();
  const int Ho = 2 * H, Wo = 2 * W;
  const long long total = static_cast<long long>(B) * Ho * Wo * C8;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(idx % C8);
    long long t = idx / C8;
    const int ox = static_cast<int>(t % Wo);
    t /= Wo;
    const int oy = static_cast<int>(t % Ho);
    const int b = static_cast<int>(t / Ho);
    const bf16x8 v =
        *reinterpret_cast<const bf16x8*>(x + ((static_cast<long long>(b) * H + (oy >> 1)) * W + (ox >> 1)) * ldx + cv * 8);
    *reinterpret_cast<bf16x8*>(y + ((static_cast<long long>(b) * Ho + oy) * Wo + ox) * ldy + cv * 8) = v;
  }
}

// ------------------------------------------------------------------ add / copy on channel slices
template <bool kAdd>
__global__ void addcopy_kernel(const __nv_bfloat16* __restrict__ a, int lda, const __nv_bfloat16* __restrict__ b,
                               int ldb, __nv_bfloat16* __restrict__ y, int ldy, long long npix, int C8) {
  pdl_prologue();
  const long long total = npix * C8;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(idx % C8);
    const long long pix = idx / C8;
    bf16x8 va = *reinterpret_cast<const bf16x8*>(a + pix * lda + cv * 8);
    if constexpr (kAdd) {
      const bf16x8 vb = *reinterpret_cast<const bf16x8*>(b + pix * ldb + cv * 8);
      float fa[8], fb[8];
      unpack8(va, fa);
      unpack8(vb, fb);
#pragma unroll
      for (int i = 0; i < 8; ++i) fa[i] += fb[i];
      va = pack8(fa);
    }
    *reinterpret_cast<bf16x8*>(y + pix * ldy + cv * 8) = va;
  }
}

// ------------------------------------------------------------------ CUDA-core conv (test cross-check only)
__global__ void conv_ref_kernel(cft_conv_args a, int Ho, int Wo, int cin_p, int kw) {
  pdl_prologue();
  const long long total = static_cast<long long>(a.B) * Ho * Wo * a.Cout;
  const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(a.x);
  const __nv_bfloat16* w = reinterpret_cast<const __nv_bfloat16*>(a.w);
  const int pad = a.k / 2, padw = kw / 2, taps = a.k * kw;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(idx % a.Cout);
    long long pix = idx / a.Cout;
    const int ox = static_cast<int>(pix % Wo);
    long long t = pix / Wo;
    const int oy = static_cast<int>(t % Ho);
    const int b = static_cast<int>(t / Ho);
    float acc = 0.f;
    for (int ky = 0; ky < a.k; ++ky)
      for (int kx = 0; kx < kw; ++kx) {
        const int iy = oy * a.stride + ky - pad, ix = ox * a.stride + kx - padw;
        if (iy < 0 || iy >= a.H || ix < 0 || ix >= a.W) continue;
        const __nv_bfloat16* xp = x + ((static_cast<long long>(b) * a.H + iy) * a.W + ix) * a.ldx + a.x_coff;
        const __nv_bfloat16* wp = w + (static_cast<long long>(n) * taps + ky * kw + kx) * cin_p;
        for (int c = 0; c < a.Cin; ++c) acc += __bfloat162float(xp[c]) * __bfloat162float(wp[c]);
      }
    if (a.bias) acc += a.bias[n];
    acc = apply_act(acc, a.act);
    if (a.out_dtype == CFT_DT_F32) {
      if (a.res) acc += reinterpret_cast<const float*>(a.res)[pix * a.ldr + a.r_coff + n];
      reinterpret_cast<float*>(a.y)[pix * a.ldy + a.y_coff + n] = acc;
    } else {
      if (a.res) acc += __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(a.res)[pix * a.ldr + a.r_coff + n]);
      reinterpret_cast<__nv_bfloat16*>(a.y)[pix * a.ldy + a.y_coff + n] = __float2bfloat16_rn(acc);
    }
  }
}

}  // namespace

using namespace cft;

extern "C" int cft_focus_gather(const void* img, int in_dtype, int B, int H, int W, long long batch_stride, int layout,
                                void* y, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  CFT_REQUIRE(img && y, "cft_focus_gather: null pointer");
  CFT_REQUIRE(B > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "cft_focus_gather: H, W must be even");
  CFT_REQUIRE(reinterpret_cast<uintptr_t>(img) % 8 == 0 && reinterpret_cast<uintptr_t>(y) % 16 == 0,
              "cft_focus_gather: misaligned pointer");
  CFT_REQUIRE(batch_stride >= 3LL * H * W && batch_stride % 2 == 0, "cft_focus_gather: bad batch stride");
  const long long total = static_cast<long long>(B) * (H / 2) * (W / 2);
  CFT_REQUIRE(layout == 0 || layout == 1, "cft_focus_gather: bad layout %d", layout);
  LaunchScope ls(CFT_K_FOCUS, stream);
  const int grid = grid_for(layout == 0 ? total : total * 4, kThreads);
  __nv_bfloat16* yo = reinterpret_cast<__nv_bfloat16*>(y);
#define CFT_FOCUS_LAUNCH(T)                                                                                        \
  do {                                                                                                             \
    if (layout == 0)                                                                                               \
      cft::launch(focus_gather_kernel<T, 0>, dim3(grid), dim3(kThreads), 0, stream, reinterpret_cast<const T*>(img), yo, B, H, W, batch_stride); \
    else                                                                                                           \
      cft::launch(focus_gather_kernel<T, 1>, dim3(grid), dim3(kThreads), 0, stream, reinterpret_cast<const T*>(img), yo, B, H, W, batch_stride); \
  } while (0)
  if (in_dtype == CFT_DT_F32) CFT_FOCUS_LAUNCH(float);
  else if (in_dtype == CFT_DT_BF16) CFT_FOCUS_LAUNCH(__nv_bfloat16);
  else if (in_dtype == CFT_DT_U8) CFT_FOCUS_LAUNCH(uint8_t);
  else return fail_arg("cft_focus_gather: bad in_dtype %d", in_dtype);
#undef CFT_FOCUS_LAUNCH
  return ls.finish("cft_focus_gather launch");
}

static int check_slice(const char* fn, const void* p, int ld, int coff, int C) {
  if (!p) return fail_arg("%s: null pointer", fn);
  if (C <= 0 || C % 8 || ld % 8 || coff % 8 || coff + C > ld)
    return fail_arg("%s: bad channel slice (C %d ld %d coff %d; need multiples of 8)", fn, C, ld, coff);
  if (reinterpret_cast<uintptr_t>(p) % 16) return fail_arg("%s: pointer not 16-byte aligned", fn);
  return CFT_OK;
}

extern "C" int cft_maxpool_s1(const void* x, int ldx, int x_coff, void* y, int ldy, int y_coff, int B, int H, int W,
                              int C, int k, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  int rc;
  if ((rc = check_slice("cft_maxpool_s1", x, ldx, x_coff, C))) return rc;
  if ((rc = check_slice("cft_maxpool_s1", y, ldy, y_coff, C))) return rc;
  CFT_REQUIRE(k >= 1 && (k & 1) && B > 0 && H > 0 && W > 0, "cft_maxpool_s1: bad k/shape");
  const long long total = static_cast<long long>(B) * H * W * (C / 8);
  LaunchScope ls(CFT_K_MAXPOOL, stream);
  cft::launch(maxpool_s1_kernel, dim3(grid_for(total, kThreads)), dim3(kThreads), 0, stream, 
      reinterpret_cast<const __nv_bfloat16*>(x) + x_coff, ldx, reinterpret_cast<__nv_bfloat16*>(y) + y_coff, ldy, B, H,
      W, C / 8, k);
  return ls.finish("cft_maxpool_s1 launch");
}


extern "C" int cft_maxpool_cascade3(const void* x, int ldx, int x_coff, void* y, int ldy, int y_coff0, int y_coff1,
                                    int y_coff2, int B, int H, int W, int C, int k0, int k1, int k2, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  int rc;
  if ((rc = check_slice("cft_maxpool_cascade3", x, ldx, x_coff, C))) return rc;
  if ((rc = check_slice("cft_maxpool_cascade3", y, ldy, y_coff0, C))) return rc;
  if ((rc = check_slice("cft_maxpool_cascade3", y, ldy, y_coff1, C))) return rc;
  if ((rc = check_slice("cft_maxpool_cascade3", y, ldy, y_coff2, C))) return rc;
  CFT_REQUIRE(C % 16 == 0 && (k0 & 1) && (k1 & 1) && (k2 & 1) && k0 > 0 && k1 > 0 && k2 > 0,
              "cft_maxpool_cascade3: C must be a multiple of 16 and the windows odd");
  CFT_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0, "cft_maxpool_cascade3: bad shape");
  const size_t smem = static_cast<size_t>(H) * W * 64;
  CFT_REQUIRE(smem <= 200 * 1024, "cft_maxpool_cascade3: plane %dx%d too large for the smem-staged kernel", H, W);
  static bool attr = false;
  if (!attr) {
    rc = check_cuda(cudaFuncSetAttribute(maxpool_cascade_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024),
                    "cudaFuncSetAttribute(maxpool_cascade)");
    if (rc) return rc;
    attr = true;
  }
  int threads = H * W;
  threads = threads > 512 ? 512 : ((threads + 31) / 32 * 32);
  dim3 grid(C / 16, B);
  LaunchScope ls(CFT_K_MAXPOOL, stream);
  __nv_bfloat16* yb = reinterpret_cast<__nv_bfloat16*>(y);
  cft::launch(maxpool_cascade_kernel, dim3(grid), dim3(threads), smem, stream, reinterpret_cast<const __nv_bfloat16*>(x) + x_coff, ldx,
                                                          yb + y_coff0, yb + y_coff1, yb + y_coff2, ldy, H, W, k0, k1, k2);
  return ls.finish("cft_maxpool_cascade3 launch");
}

extern "C" int cft_upsample2x(const void* x, int ldx, int x_coff, void* y, int ldy, int y_coff, int B, int H, int W,
                              int C, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  int rc;
  if ((rc = check_slice("cft_upsample2x", x, ldx, x_coff, C))) return rc;
  if ((rc = check_slice("cft_upsample2x", y, ldy, y_coff, C))) return rc;
  CFT_REQUIRE(B > 0 && H > 0 && W > 0, "cft_upsample2x: empty shape");
  const long long total = static_cast<long long>(B) * H * W * 4 * (C / 8);
  LaunchScope ls(CFT_K_UPSAMPLE, stream);
  cft::launch(upsample2x_kernel, dim3(grid_for(total, kThreads)), dim3(kThreads), 0, stream, 
      reinterpret_cast<const __nv_bfloat16*>(x) + x_coff, ldx, reinterpret_cast<__nv_bfloat16*>(y) + y_coff, ldy, B, H,
      W, C / 8);
  return ls.finish("cft_upsample2x launch");
}

extern "C" int cft_add(const void* a, int lda, int a_coff, const void* b, int ldb, int b_coff, void* y, int ldy,
                       int y_coff, long long npix, int C, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  int rc;
  if ((rc = check_slice("cft_add", a, lda, a_coff, C))) return rc;
  if ((rc = check_slice("cft_add", b, ldb, b_coff, C))) return rc;
  if ((rc = check_slice("cft_add", y, ldy, y_coff, C))) return rc;
  CFT_REQUIRE(npix > 0, "cft_add: empty");
  LaunchScope ls(CFT_K_ADD, stream);
  cft::launch(addcopy_kernel<true>, dim3(grid_for(npix * (C / 8), kThreads)), dim3(kThreads), 0, stream, 
      reinterpret_cast<const __nv_bfloat16*>(a) + a_coff, lda, reinterpret_cast<const __nv_bfloat16*>(b) + b_coff, ldb,
      reinterpret_cast<__nv_bfloat16*>(y) + y_coff, ldy, npix, C / 8);
  return ls.finish("cft_add launch");
}

extern "C" int cft_copy(const void* x, int ldx, int x_coff, void* y, int ldy, int y_coff, long long npix, int C,
                        void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  int rc;
  if ((rc = check_slice("cft_copy", x, ldx, x_coff, C))) return rc;
  if ((rc = check_slice("cft_copy", y, ldy, y_coff, C))) return rc;
  CFT_REQUIRE(npix > 0, "cft_copy: empty");
  LaunchScope ls(CFT_K_COPY, stream);
  cft::launch(addcopy_kernel<false>, dim3(grid_for(npix * (C / 8), kThreads)), dim3(kThreads), 0, stream, 
      reinterpret_cast<const __nv_bfloat16*>(x) + x_coff, ldx, nullptr, 0,
      reinterpret_cast<__nv_bfloat16*>(y) + y_coff, ldy, npix, C / 8);
  return ls.finish("cft_copy launch");
}

extern "C" int cft_conv2d_ref(const cft_conv_args* a, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  CFT_REQUIRE(a && a->x && a->w && a->y, "cft_conv2d_ref: null pointer");
  CFT_REQUIRE((a->k == 1 || a->k == 3) && (a->stride == 1 || a->stride == 2), "cft_conv2d_ref: bad k/stride");
  const int Ho = (a->H + a->stride - 1) / a->stride, Wo = (a->W + a->stride - 1) / a->stride;
  const int cin_p = (a->Cin + 7) / 8 * 8;
  const long long total = static_cast<long long>(a->B) * Ho * Wo * a->Cout;
  LaunchScope ls(CFT_K_CONV_REF, stream);
  const int kw = a->kw > 0 ? a->kw : a->k;
  cft::launch(conv_ref_kernel, dim3(grid_for(total, kThreads, 32)), dim3(kThreads), 0, stream, *a, Ho, Wo, cin_p, kw);
  return ls.finish("cft_conv2d_ref launch");
}
