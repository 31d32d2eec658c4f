// ndns.cu - NDNS pre/post-processing either side of the model forward.
//
//   s5b200_stft      : stft_splitter, train_helpers.py:1381-1396 (jax.scipy.signal.stft: boxcar, nperseg = nfft = 512,
//                      hop 128, one-sided, boundary='zeros', padded=True, scaling='spectrum' = 1 / 512) -> magnitude, angle
//                      and, fused, the model input  mag - 0.0007  in the model's (B, L, 257) layout (:1570-1573)
//   s5b200_istft     : mask apply  mag * (1 + mask)  (:1605-1606) + stft_mixer, :1399-1412 (un-normalised inverse real
//                      DFT per frame, overlap-add, division by the number of covering frames, 256 samples trimmed)
//   s5b200_si_snr    : si_snr_jax, :16-53
//   s5b200_mse_clips : mean((a - b)^2) per clip, the magnitude term of the loss (:1609-1611)
//
// All of it is HBM-bound streaming work: 61 MB of audio against 2 x 123 MB of spectra per 32-clip batch.  A 512-point
// real transform is one 256-point complex Stockham radix-2 FFT per warp in shared memory (x[2n] + i x[2n+1]) plus the
// split / merge pass; twiddles come from one sincospif table per CTA.
#include <cmath>

#include "common.cuh"

namespace s5 {

constexpr int kNfft = 512, kHop = 128, kBins = 257;
constexpr int kFB = 32;            // frames (STFT) / hops (iSTFT) per CTA
constexpr int kNdThreads = 256;

__device__ __forceinline__ float2 cmul2(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// 256-point complex FFT, one warp, ping-pong buffers in shared memory: RADIX-4 Stockham, 4 stages (the result ends in `a`), two
// butterflies per lane and stage.  Stage s (Ns = 4^s): v_m = in[j + 64 m] * exp(-2 pi i m r / (4 Ns)), r = j mod Ns, a 4-point DFT,
// out[4 (j - r) + r + m Ns].  The first version was radix 2 (8 stages): ncu showed both transforms between the issue limit (71 %)
// and the shared-memory limit (62-70 % of the wavefront peak); radix 4 halves the passes over shared memory (88 instead of 160
// accesses per lane and transform) and the butterfly instructions.
// Twiddles: ONE TABLE PER STAGE at offset 4^s - 1, entry (m - 1) Ns + r, so the lanes of a warp read consecutive (or repeated)
// words - indexing one table exp(-2 pi i k / 512) with a stride put up to 16 lanes on one bank (ncu: the iSTFT at 79 % of the
// shared-memory wavefront peak with 36.6 M bank conflicts).  INV: conjugated twiddles and butterflies (un-normalised inverse).
// kTwFull: exp(-2 pi i k / 512), k < 256, for the split / merge passes.
constexpr int kTwEntries = 511, kTwFull = 255;
template <bool INV>
__device__ __forceinline__ void fft256_warp(float2* a, float2* b, const float2* tw, int lane) {
  float2* in = a;
  float2* out = b;
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int Ns = 1 << (2 * s);
    const float2* t = tw + (Ns - 1);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int j = lane + 32 * q, r = j & (Ns - 1);
      const float2 v0 = in[j];
      float2 v1 = in[j + 64], v2 = in[j + 128], v3 = in[j + 192];
      if (s > 0) {                                             // stage 0: all twiddles are 1
        float2 w1 = t[r], w2 = t[Ns + r], w3 = t[2 * Ns + r];
        if (INV) { w1.y = -w1.y; w2.y = -w2.y; w3.y = -w3.y; }
        v1 = cmul2(v1, w1); v2 = cmul2(v2, w2); v3 = cmul2(v3, w3);
      }
      const float2 A = make_float2(v0.x + v2.x, v0.y + v2.y), B = make_float2(v0.x - v2.x, v0.y - v2.y);
      const float2 Cs = make_float2(v1.x + v3.x, v1.y + v3.y), D = make_float2(v1.x - v3.x, v1.y - v3.y);
      const float2 y0 = make_float2(A.x + Cs.x, A.y + Cs.y), y2 = make_float2(A.x - Cs.x, A.y - Cs.y);
      // forward: y1 = B - i D, y3 = B + i D; inverse: the other way round
      const float2 bm = make_float2(B.x + D.y, B.y - D.x), bp = make_float2(B.x - D.y, B.y + D.x);
      const float2 y1 = INV ? bp : bm, y3 = INV ? bm : bp;
      const int j0 = ((j - r) << 2) + r;
      if (s == 0) {                                            // out[4 j .. 4 j + 3]: two 16-byte stores
        *reinterpret_cast<float4*>(out + j0) = make_float4(y0.x, y0.y, y1.x, y1.y);
        *reinterpret_cast<float4*>(out + j0 + 2) = make_float4(y2.x, y2.y, y3.x, y3.y);
      } else {
        out[j0] = y0;
        out[j0 + Ns] = y1;
        out[j0 + 2 * Ns] = y2;
        out[j0 + 3 * Ns] = y3;
      }
    }
    __syncwarp();
    float2* tmp = in; in = out; out = tmp;
  }
}

__device__ __forceinline__ void make_twiddles(float2* tw) {
  for (int i = threadIdx.x; i < kTwEntries; i += blockDim.x) {
    float x;                                                   // entry = exp(-i pi x)
    if (i < kTwFull) {
      const int s = (31 - __clz(i + 1)) >> 1, Ns = 1 << (2 * s), off = i - (Ns - 1);
      const int m = off / Ns + 1, r = off - (m - 1) * Ns;
      x = (float)(m * r) / (float)(2 * Ns);                    // exp(-2 pi i m r / (4 Ns)); exact in fp32
    } else {
      x = (float)(i - kTwFull) * (1.0f / 256.0f);             // exp(-2 pi i k / 512)
    }
    float sn, cs;
    sincospif(x, &sn, &cs);
    tw[i] = make_float2(cs, -sn);
  }
}

// ---------------------------------------------------------------------------------------------- STFT
struct StftParams {
  const float* audio; int B, T, L;
  float* mag; float* phase; int blf;   // blf: outputs (B, L, 257); else the reference's (B, 257, L)
  float* model_in; float mag_sub;      // (B, L, 257) = mag - mag_sub
};

// BLF (the model's (B, L, 257) layout): a frame's 257 bins are one contiguous row, so the warp that transformed the frame writes
// magnitude, angle and model input straight from the split pass - no staging tile, 52 KB of shared memory per CTA, four CTAs
// (32 warps) per SM instead of one (the first version staged every layout: 118 KB, 8 warps per SM, latency bound).  The
// reference's (B, 257, L) layout still goes through the [32][257] staging tiles for its transposed write-out.
template <bool BLF>
__global__ void __launch_bounds__(kNdThreads, BLF ? 4 : 1) stft_kernel(StftParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int kSeg = (kFB - 1) * kHop + kNfft;              // samples the CTA's frames cover
  float* seg = reinterpret_cast<float*>(smem);
  float2* tw = reinterpret_cast<float2*>(seg + kSeg);
  float2* work = tw + 512;                                    // [8 warps][2][256]
  float* omag = reinterpret_cast<float*>(work + 8 * 512);     // [kFB][257] (not BLF only)
  float* oph = omag + kFB * kBins;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y, k0 = blockIdx.x * kFB, nf = min(kFB, p.L - k0);
  make_twiddles(tw);
  const float* au = p.audio + (size_t)b * p.T;
  const long long t0 = (long long)k0 * kHop - kNfft / 2;      // audio index of the segment's first sample
  for (int i = tid; i < kSeg; i += kNdThreads) {
    const long long t = t0 + i;
    seg[i] = (t >= 0 && t < p.T) ? au[t] : 0.f;
  }
  __syncthreads();
  const size_t clip_off = (size_t)b * p.L * kBins;
  float2* za = work + warp * 512;
  float2* zb = za + 256;
  for (int f = warp; f < nf; f += 8) {
    const float2* x2 = reinterpret_cast<const float2*>(seg + f * kHop);   // z[n] = x[2n] + i x[2n+1]
#pragma unroll
    for (int q = 0; q < 8; ++q) za[lane + 32 * q] = x2[lane + 32 * q];
    __syncwarp();
    fft256_warp<false>(za, zb, tw, lane);
    // split: X[k] = E + W^k O,  E = (Z[k] + conj Z[256 - k]) / 2,  O = -i (Z[k] - conj Z[256 - k]) / 2;  scaled by 1 / 512
    for (int k = lane; k < kBins; k += 32) {
      const float2 zk = za[k & 255], zc = za[(256 - k) & 255];
      const float2 E = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y - zc.y));
      const float2 O = make_float2(0.5f * (zk.y + zc.y), -0.5f * (zk.x - zc.x));
      const float2 w = k < 256 ? tw[kTwFull + k] : make_float2(-1.f, 0.f);
      const float2 wo = cmul2(w, O);
      const float re = (E.x + wo.x) * (1.0f / kNfft), im = (E.y + wo.y) * (1.0f / kNfft);
      const float mg = hypotf(re, im);
      if (BLF) {
        const size_t o = clip_off + (size_t)(k0 + f) * kBins + k;
        if (p.mag) p.mag[o] = mg;
        if (p.phase) p.phase[o] = atan2f(im, re);
        if (p.model_in) p.model_in[o] = mg - p.mag_sub;
      } else {
        omag[f * kBins + k] = mg;
        oph[f * kBins + k] = atan2f(im, re);
      }
    }
    __syncwarp();
  }
  if (BLF) return;
  __syncthreads();
  // ---- write-out: the model input is contiguous per frame; BFL is 32 consecutive frames per bin
  if (p.model_in) {
    float* dst = p.model_in + clip_off + (size_t)k0 * kBins;
    for (int i = tid; i < nf * kBins; i += kNdThreads) dst[i] = omag[i] - p.mag_sub;
  }
  {
    for (int k = warp; k < kBins; k += 8) {
      if (lane < nf) {
        const size_t o = clip_off + (size_t)k * p.L + k0 + lane;
        if (p.mag) p.mag[o] = omag[lane * kBins + k];
        if (p.phase) p.phase[o] = oph[lane * kBins + k];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- iSTFT
struct IstftParams {
  const float* mag; const float* phase; int blf;
  const float* mask;                   // model output (B, L, 257), nullable: cleaned_mag = mag * (1 + mask)
  int B, L;
  float* audio;                        // (B, (L - 1) * 128)
  float* cleaned_mag;                  // nullable, layout of mag
};

// 29 hops of output per CTA = 32 frames (rows) = two rounds of 16 warps; a warp's second FFT buffer is the spectrum row it has
// just consumed, so the CTA needs 100 KB and two CTAs (32 warps) share an SM (first version: 35 rows over 8 warps, 16 warps per SM).
constexpr int kIB = 29;            // hops per CTA (iSTFT)
constexpr int kIsThreads = 512;
__global__ void __launch_bounds__(kIsThreads, 2) istft_kernel(IstftParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int kFr = kIB + 3;                                // frames that touch the CTA's 29 hops of output
  constexpr int kRow = 260;                                   // float2 per spectrum row (257 bins, padded); 512 floats of time fit too
  float2* S = reinterpret_cast<float2*>(smem);                // [kFr][kRow] spectra, then (in place) the frames' time samples
  float2* tw = S + kFr * kRow;
  float2* work = tw + 512;                                    // [16][256]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y, h0 = blockIdx.x * kIB;            // output hops h0 .. h0 + 28 (padded coordinates)
  const int f_lo = h0 - 3;                                    // frame of S row 0
  make_twiddles(tw);
  const size_t clip_off = (size_t)b * p.L * kBins;
  // ---- load: magnitude (* (1 + mask)), angle -> complex spectrum rows
  if (p.blf) {
    // four elements per thread and pass, all of their loads issued before the first sincos (12 requests in flight per thread)
    constexpr int kU = 4;
    const float* __restrict__ gmag = p.mag;
    const float* __restrict__ gph = p.phase;
    const float* __restrict__ gmask = p.mask;
    for (int i0 = tid; i0 < kFr * kBins; i0 += kIsThreads * kU) {
      float m[kU], ph[kU], mk[kU];
      int f[kU], sidx[kU];
      size_t o[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int i = i0 + u * kIsThreads, r = i / kBins, k = i - r * kBins;
        f[u] = i < kFr * kBins ? f_lo + r : -1;
        sidx[u] = r * kRow + k;
        o[u] = clip_off + (size_t)max(f[u], 0) * kBins + k;
        const bool ok = f[u] >= 0 && f[u] < p.L;
        m[u] = ok ? gmag[o[u]] : 0.f;
        ph[u] = ok ? gph[o[u]] : 0.f;
        mk[u] = ok && gmask ? gmask[o[u]] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        if (f[u] < 0 || f[u] >= p.L) continue;
        float mm = m[u];
        if (gmask) mm *= 1.0f + mk[u];
        // this CTA owns the cleaned magnitude of the frames that START in its hops (f >= h0)
        if (p.cleaned_mag && f[u] >= h0 && f[u] < h0 + kIB) p.cleaned_mag[o[u]] = mm;
        float sn, cs;
        sincosf(ph[u], &sn, &cs);
        S[sidx[u]] = make_float2(mm * cs, mm * sn);
      }
    }
  } else {
    for (int i = tid; i < kBins * 64; i += kIsThreads) {      // lanes along frames (contiguous in the (B, 257, L) layout)
      const int k = i >> 6, r = i & 63, f = f_lo + r;
      if (r >= kFr || f < 0 || f >= p.L) continue;
      const size_t o = clip_off + (size_t)k * p.L + f;
      float m = p.mag[o];
      if (p.mask) m *= 1.0f + p.mask[clip_off + (size_t)f * kBins + k];
      if (p.cleaned_mag && f >= h0 && f < h0 + kIB) p.cleaned_mag[o] = m;
      float s, c;
      sincosf(p.phase[o], &s, &c);
      S[r * kRow + k] = make_float2(m * c, m * s);
    }
  }
  __syncthreads();
  // ---- per frame: merge to the 256-point spectrum of x[2n] + i x[2n+1], inverse FFT, time samples back into the row
  float2* za = work + warp * 256;
  for (int r = warp; r < kFr; r += kIsThreads / 32) {
    const int f = f_lo + r;
    if (f < 0 || f >= p.L) continue;                          // warp-uniform
    float2* row = S + r * kRow;
    float2* zb = row;                                         // the row is dead once it has been merged into za
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int k = lane + 32 * q;
      float2 xk = row[k], xc = row[256 - k];
      if (k == 0) { xk.y = 0.f; xc.y = 0.f; }                 // irfft ignores the imaginary parts of DC and Nyquist
      const float2 E = make_float2(0.5f * (xk.x + xc.x), 0.5f * (xk.y - xc.y));
      const float2 D = make_float2(0.5f * (xk.x - xc.x), 0.5f * (xk.y + xc.y));
      const float2 w = make_float2(tw[kTwFull + k].x, -tw[kTwFull + k].y);   // exp(+2 pi i k / 512)
      const float2 O = cmul2(D, w);
      za[k] = make_float2(E.x - O.y, E.y + O.x);              // Z = E + i O
    }
    __syncwarp();
    fft256_warp<true>(za, zb, tw, lane);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float2 z = za[lane + 32 * q];
      row[lane + 32 * q] = make_float2(2.f * z.x, 2.f * z.y);   // x[2n], x[2n + 1] of irfft * 512
    }
    __syncwarp();
  }
  __syncthreads();
  // ---- overlap-add in frame order, divide by the number of covering frames, trim 256 samples either side
  const int out_len = (p.L - 1) * kHop;
  const float* Sf = reinterpret_cast<const float*>(S);
  for (int i = tid; i < kIB * kHop; i += kIsThreads) {
    const int h = h0 + (i >> 7), within = i & 127;
    const long long t = (long long)h * kHop + within - kNfft / 2;   // output index
    if (t < 0 || t >= out_len) continue;
    float acc = 0.f;
    int cnt = 0;
#pragma unroll
    for (int d = 3; d >= 0; --d) {                            // frames h - 3 .. h, increasing
      const int f = h - d;
      if (f < 0 || f >= p.L) continue;
      acc += Sf[(size_t)(f - f_lo) * (2 * kRow) + d * kHop + within];
      ++cnt;
    }
    p.audio[(size_t)b * out_len + t] = acc / (float)cnt;
  }
}

// ---------------------------------------------------------------------------------------------- per-clip reductions
// partials[b][chunk][5] (double): sum t, sum e, sum t^2, sum e^2, sum t e over one chunk; fixed-order finalisation.
constexpr int kRedChunk = 16384;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int NS>
__device__ __forceinline__ void block_sums(double (&v)[NS], double* out) {
  __shared__ double sh[NS][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    v[s] = warp_sum(v[s]);
    if (lane == 0) sh[s][warp] = v[s];
  }
  __syncthreads();
  if (threadIdx.x < NS) {
    double a = 0;
    for (int w = 0; w < 8; ++w) a += sh[threadIdx.x][w];
    out[threadIdx.x] = a;
  }
}

__global__ void __launch_bounds__(kNdThreads) snr_partial_kernel(const float* target, const float* est, long long T, int nchunk, double* partials) {
  const int b = blockIdx.y, c = blockIdx.x;
  const long long i0 = (long long)c * kRedChunk, i1 = min(T, i0 + kRedChunk);
  const float* t = target + (size_t)b * T;
  const float* e = est + (size_t)b * T;
  double v[5] = {0, 0, 0, 0, 0};
  for (long long i = i0 + threadIdx.x; i < i1; i += kNdThreads) {
    const double a = t[i], x = e[i];
    v[0] += a; v[1] += x; v[2] += a * a; v[3] += x * x; v[4] += a * x;
  }
  block_sums<5>(v, partials + ((size_t)b * nchunk + c) * 5);
}

__global__ void snr_final_kernel(const double* partials, long long T, int nchunk, int B, float* out_db) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double s[5] = {0, 0, 0, 0, 0};
  for (int c = 0; c < nchunk; ++c)
    for (int j = 0; j < 5; ++j) s[j] += partials[((size_t)b * nchunk + c) * 5 + j];
  // zero-mean signals: dot = sum te - T mt me, |s_t|^2 = sum t^2 - T mt^2, |s_e|^2 likewise;
  // |proj|^2 = dot^2 / |s_t|^2, |noise|^2 = |s_e|^2 - dot^2 / |s_t|^2   (train_helpers.py:39-53)
  const double n = (double)T, mt = s[0] / n, me = s[1] / n;
  const double dot = s[4] - n * mt * me, tt = s[2] - n * mt * mt, ee = s[3] - n * me * me;
  const double proj2 = dot * dot / tt, noise2 = fmax(ee - proj2, 0.0);
  out_db[b] = (float)(10.0 * log10(proj2 / (noise2 + 1e-8) + 1e-8));
}

__global__ void __launch_bounds__(kNdThreads) mse_partial_kernel(const float* a, const float* bb, long long n, int nchunk, double* partials) {
  const int b = blockIdx.y, c = blockIdx.x;
  const long long i0 = (long long)c * kRedChunk, i1 = min(n, i0 + kRedChunk);
  double v[1] = {0};
  for (long long i = i0 + threadIdx.x; i < i1; i += kNdThreads) {
    const double d = (double)a[(size_t)b * n + i] - (double)bb[(size_t)b * n + i];
    v[0] += d * d;
  }
  block_sums<1>(v, partials + (size_t)b * nchunk + c);
}
__global__ void mse_final_kernel(const double* partials, long long n, int nchunk, int B, float* out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double s = 0;
  for (int c = 0; c < nchunk; ++c) s += partials[(size_t)b * nchunk + c];
  out[b] = (float)(s / (double)n);
}

static int red_chunks(long long n) { return (int)((n + kRedChunk - 1) / kRedChunk); }

}  // namespace s5

using namespace s5;

extern "C" int s5b200_stft_frames(int T) {
  if (T <= 0) return 0;
  const long long ext = (long long)T + kNfft;
  const long long pad = (kHop - (ext - kNfft) % kHop) % kHop;
  return (int)((ext + pad - kNfft) / kHop + 1);
}

extern "C" int s5b200_stft(const float* audio, int B, int T, float* mag, float* phase, int layout_blf, float* model_in, float mag_sub,
                           void* stream) {
  if (!audio || (!mag && !phase && !model_in)) S5_FAIL(S5_E_BADARG, "stft: null pointer");
  if (B <= 0 || T <= 0 || B > 65535) S5_FAIL(S5_E_BADSHAPE, "stft: B=%d T=%d outside the supported range", B, T);
  StftParams p{audio, B, T, s5b200_stft_frames(T), mag, phase, layout_blf, model_in, mag_sub};
  const size_t smem = (size_t)((kFB - 1) * kHop + kNfft) * 4 + 512 * 8 + 8 * 512 * 8 + (layout_blf ? 0 : 2 * (size_t)kFB * kBins * 4);
  auto kern = layout_blf ? stft_kernel<true> : stft_kernel<false>;
  S5_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope prof(S5_K_OTHER, (cudaStream_t)stream);
  kern<<<dim3((p.L + kFB - 1) / kFB, B), kNdThreads, smem, (cudaStream_t)stream>>>(p);
  S5_LAUNCH_CHECK();
  return S5_OK;
}

extern "C" int s5b200_istft(const float* mag, const float* phase, int layout_blf, const float* mask_blf, int B, int L, float* audio_out,
                            float* cleaned_mag_out, void* stream) {
  if (!mag || !phase || !audio_out) S5_FAIL(S5_E_BADARG, "istft: null pointer");
  if (B <= 0 || L < 2 || B > 65535) S5_FAIL(S5_E_BADSHAPE, "istft: B=%d L=%d outside the supported range", B, L);
  // a CTA also reads the three frames before its own hops, which another CTA owns and would overwrite
  if (cleaned_mag_out && (cleaned_mag_out == mag || cleaned_mag_out == phase || cleaned_mag_out == mask_blf))
    S5_FAIL(S5_E_BADARG, "istft: cleaned_mag_out must not alias an input");
  IstftParams p{mag, phase, layout_blf, mask_blf, B, L, audio_out, cleaned_mag_out};
  const size_t smem = (size_t)(kIB + 3) * 260 * 8 + 512 * 8 + (kIsThreads / 32) * 256 * 8;
  S5_CUDA(cudaFuncSetAttribute(istft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // hops 0 .. L + 2 carry output samples (frame L - 1 ends at hop L + 2)
  ProfScope prof(S5_K_OTHER, (cudaStream_t)stream);
  istft_kernel<<<dim3((L + 3 + kIB - 1) / kIB, B), kIsThreads, smem, (cudaStream_t)stream>>>(p);
  S5_LAUNCH_CHECK();
  return S5_OK;
}

extern "C" size_t s5b200_reduce_workspace_bytes(int B, long long n_per_clip) {
  if (B <= 0 || n_per_clip <= 0) return 0;
  return (size_t)B * red_chunks(n_per_clip) * 5 * sizeof(double);
}

extern "C" int s5b200_si_snr(const float* target, const float* estimate, int B, long long T, float* out_db, void* workspace,
                             size_t workspace_bytes, void* stream) {
  if (!target || !estimate || !out_db || !workspace) S5_FAIL(S5_E_BADARG, "si_snr: null pointer");
  if (B <= 0 || T <= 0 || B > 65535) S5_FAIL(S5_E_BADSHAPE, "si_snr: B=%d T=%lld outside the supported range", B, T);
  if (workspace_bytes < s5b200_reduce_workspace_bytes(B, T)) S5_FAIL(S5_E_BADARG, "si_snr: workspace too small");
  const int nchunk = red_chunks(T);
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(S5_K_OTHER, st);
  snr_partial_kernel<<<dim3(nchunk, B), kNdThreads, 0, st>>>(target, estimate, T, nchunk, reinterpret_cast<double*>(workspace));
  S5_LAUNCH_CHECK();
  snr_final_kernel<<<(B + 127) / 128, 128, 0, st>>>(reinterpret_cast<const double*>(workspace), T, nchunk, B, out_db);
  S5_LAUNCH_CHECK();
  return S5_OK;
}

extern "C" int s5b200_mse_clips(const float* a, const float* b, int B, long long n_per_clip, float* out, void* workspace,
                                size_t workspace_bytes, void* stream) {
  if (!a || !b || !out || !workspace) S5_FAIL(S5_E_BADARG, "mse_clips: null pointer");
  if (B <= 0 || n_per_clip <= 0 || B > 65535) S5_FAIL(S5_E_BADSHAPE, "mse_clips: B=%d n=%lld outside the supported range", B, n_per_clip);
  if (workspace_bytes < s5b200_reduce_workspace_bytes(B, n_per_clip)) S5_FAIL(S5_E_BADARG, "mse_clips: workspace too small");
  const int nchunk = red_chunks(n_per_clip);
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(S5_K_OTHER, st);
  mse_partial_kernel<<<dim3(nchunk, B), kNdThreads, 0, st>>>(a, b, n_per_clip, nchunk, reinterpret_cast<double*>(workspace));
  S5_LAUNCH_CHECK();
  mse_final_kernel<<<(B + 127) / 128, 128, 0, st>>>(reinterpret_cast<const double*>(workspace), n_per_clip, nchunk, B, out);
  S5_LAUNCH_CHECK();
  return S5_OK;
}
