// Batched STFT / iSTFT + amplitude compression on the device: the step either side of the sampler.
//
// Reference semantics (paths relative to /root/reference):
//   forward  VFModel._stft -> SpecsDataModule.stft = torch.stft(n_fft=510, hop=128, hann(periodic), center=True (reflect),
//            onesided, return_complex) (flowmse/data_module.py:149-170, model.py:196-199), then spec_fwd:
//            |X|^e * exp(j angle X) * factor (data_module.py:149-162), then pad_spec: zero-pad T to a multiple of 64
//            (util/other.py:83-90); evaluate.py:109-110 divides the waveform by its peak first.
//   inverse  spec_back (data_module.py:164-175) -> torch.istft(..., length) (data_module.py:172-175, model.py:190-203),
//            times the peak (evaluate.py:134-135).
// The reference does this per file on one utterance at a time with a .item() host sync for the peak; here a whole batch
// of ragged utterances goes through five launches with no host synchronisation:
//   wav_prep (reflect padding + peak via atomicMax) -> SIMT GEMM frames x windowed-DFT basis (the frame matrix is the
//   padded signal itself read with a row stride of one hop) -> spec_fwd (transpose to [F,T], 1/peak, compression, zero
//   padding); and spec_back (decompression, transpose) -> GEMM with the windowed inverse-DFT basis -> overlap-add /
//   window-envelope division.  n_fft = 510 -> 256 bins is what the backbone requires (F = 256), so it is fixed here.
#include "flowse_internal.h"

#include <cmath>

namespace flowse {

namespace {

constexpr int kNfft = 510, kHop = 128, kBins = 256, kPad = kNfft / 2;   // center=True pads n_fft/2 = 255 on both sides
constexpr double kTwoPi = 6.283185307179586476925286766559;
constexpr int kSpecLog = 1, kSpecNone = 2;        // 0 = exponent;   // transform_type (data_module.py:149-175)

// hann(510, periodic) in fp32, exactly torch.hann_window's formula evaluated in double and rounded
__device__ __forceinline__ float hann(int n) {
  return static_cast<float>(0.5 - 0.5 * cos(kTwoPi * n / kNfft));
}
// get_window (data_module.py:13-19): 'hann' or 'sqrthann' = torch.sqrt of the fp32 hann window
__device__ __forceinline__ float window_at(int n, int sqrt_window) {
  const float h = hann(n);
  return sqrt_window ? sqrtf(h) : h;
}

// fwd [510][512]: column 2f = w[k] cos(2 pi f k / 510), column 2f+1 = -w[k] sin(...)
// inv [512][512]: row 2f = a_f w[n] cos(2 pi f n / 510) / 510, row 2f+1 = -a_f w[n] sin(...) / 510, columns >= 510 zero;
//                 a_f = 1 for DC and Nyquist (f = 0, 255), else 2 (onesided C2R)
__global__ void stft_basis_kernel(float* __restrict__ fwd, float* __restrict__ inv, float* __restrict__ win, int sqrt_window) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < kNfft) win[i] = window_at(i, sqrt_window);
  if (i < kNfft * 512) {
    const int k = i / 512, col = i % 512, f = col >> 1;
    const int m = (f * k) % kNfft;                       // exact argument reduction
    const double ang = kTwoPi * m / kNfft;
    const double w = static_cast<double>(window_at(k, sqrt_window));
    fwd[i] = static_cast<float>((col & 1) ? -w * sin(ang) : w * cos(ang));
  }
  if (i < 512 * 512) {
    const int row = i / 512, n = i % 512, f = row >> 1;
    float v = 0.f;
    if (n < kNfft) {
      const int m = (f * n) % kNfft;
      const double ang = kTwoPi * m / kNfft;
      const double a = (f == 0 || f == kBins - 1) ? 1.0 : 2.0;
      const double w = static_cast<double>(window_at(n, sqrt_window));
      v = static_cast<float>(((row & 1) ? -a * w * sin(ang) : a * w * cos(ang)) / kNfft);
    }
    inv[i] = v;
  }
}

// xpad[b][i] = reflect-padded utterance (i in [0, L + 510)), zeros beyond; peak[b] = max |wav| (bits of a non-negative
// float order like unsigned integers, so atomicMax on the bit pattern is exact and order-independent)
__global__ void __launch_bounds__(256)
wav_prep_kernel(const float* __restrict__ wav, long long wav_stride, const int* __restrict__ lengths,
                float* __restrict__ xpad, long long xpad_stride, unsigned* __restrict__ peak_bits) {
  const int b = blockIdx.y;
  const int L = lengths[b];
  const float* src = wav + b * wav_stride;
  float* dst = xpad + b * xpad_stride;
  float mx = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < xpad_stride;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float v = 0.f;
    if (i < L + 2 * kPad) {
      long long j = i - kPad;
      if (j < 0) j = -j;
      else if (j >= L) j = 2LL * (L - 1) - j;
      v = src[j];
      if (i >= kPad && i < L + kPad) mx = fmaxf(mx, fabsf(v));
    }
    dst[i] = v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (peak_bits && (threadIdx.x & 31) == 0 && mx > 0.f) atomicMax(peak_bits + b, __float_as_uint(mx));
}

// S [B][Tmax][256] complex (time-major GEMM output) -> Y [B][256][Tpad] complex: 1/peak, |X|^e e^{j angle}, * factor;
// frames t >= frames(b) are the zero padding of pad_spec.  32 x 32 tiles through shared memory (both sides coalesced).
__global__ void __launch_bounds__(256)
spec_fwd_kernel(const float2* __restrict__ S, int Tmax, const int* __restrict__ lengths, const unsigned* __restrict__ peak_bits,
                float factor, float expo, int mode, float2* __restrict__ Y, int Tpad) {
  __shared__ float2 tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
  const int Tb = 1 + lengths[b] / kHop;
  const float inv_peak = peak_bits ? 1.0f / __uint_as_float(peak_bits[b]) : 1.0f;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int t = t0 + r;
    float2 v = make_float2(0.f, 0.f);
    if (t < Tb && t < Tmax) v = S[(static_cast<size_t>(b) * Tmax + t) * kBins + f0 + tx];
    tile[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int t = t0 + tx;
    if (t >= Tpad) continue;
    float2 v = tile[tx][r];
    v.x *= inv_peak; v.y *= inv_peak;
    const float mag = sqrtf(v.x * v.x + v.y * v.y);
    float2 o = make_float2(0.f, 0.f);
    if (mode == kSpecNone) {
      o = v;
    } else if (mag > 0.f) {
      // exponent: |X|^e * X / |X| * factor;  log: log(1 + |X|) * X / |X| * factor
      const float c = mode == kSpecLog ? log1pf(mag) : (expo == 0.5f ? sqrtf(mag) : (expo == 1.0f ? mag : powf(mag, expo)));
      const float g = c / mag * factor;
      o.x = v.x * g; o.y = v.y * g;
    }
    Y[(static_cast<size_t>(b) * kBins + f0 + r) * Tpad + t] = o;
  }
}

// X [B][256][Tpad] complex -> S [B][Tmax][256] complex (time-major, GEMM operand): / factor, |X|^(1/e) e^{j angle}
__global__ void __launch_bounds__(256)
spec_back_kernel(const float2* __restrict__ X, int Tpad, float factor, float expo, int mode, float2* __restrict__ S,
                 int Tmax) {
  __shared__ float2 tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float inv_factor = mode == kSpecNone ? 1.0f : 1.0f / factor;
  for (int r = ty; r < 32; r += 8) {
    const int t = t0 + tx;
    float2 o = make_float2(0.f, 0.f);
    if (t < Tpad && t < Tmax) {
      float2 v = X[(static_cast<size_t>(b) * kBins + f0 + r) * Tpad + t];
      v.x *= inv_factor; v.y *= inv_factor;
      const float mag = sqrtf(v.x * v.x + v.y * v.y);
      if (mode == kSpecNone) {
        o = v;
      } else if (mag > 0.f) {
        const float c = mode == kSpecLog ? expm1f(mag) : (expo == 0.5f ? mag * mag : (expo == 1.0f ? mag : powf(mag, 1.0f / expo)));
        const float g = c / mag;
        o.x = v.x * g; o.y = v.y * g;
      }
    }
    tile[r][tx] = o;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int t = t0 + r;
    if (t < Tmax) S[(static_cast<size_t>(b) * Tmax + t) * kBins + f0 + tx] = tile[tx][r];
  }
}

// torch.istft's overlap-add: out[s] = peak * sum_t frames[t][s + 255 - 128 t] / sum_t w^2[s + 255 - 128 t] over the frames
// that cover sample s (at most 4); frames already carry the synthesis window (folded into the basis).  ALL T frames of the
// padded spectrogram take part, as in the reference: evaluate.py:132 hands the pad_spec-padded sampler output to to_audio,
// so the frames beyond an utterance's own count contribute to (and weigh in the envelope of) its last ~3 hops.
__global__ void __launch_bounds__(256)
ola_kernel(const float* __restrict__ frames, int Tmax, const int* __restrict__ lengths, const float* __restrict__ win,
           const float* __restrict__ peak, float* __restrict__ out, long long out_stride) {
  const int b = blockIdx.y;
  const int L = lengths[b];
  const int Tb = Tmax;
  const float pk = peak ? peak[b] : 1.0f;
  for (long long s = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; s < out_stride;
       s += static_cast<long long>(gridDim.x) * blockDim.x) {
    float v = 0.f;
    if (s < L) {
      const long long q = s + kPad;                         // position in the padded (center=True) signal
      int t_hi = static_cast<int>(q / kHop);
      if (t_hi > Tb - 1) t_hi = Tb - 1;
      float acc = 0.f, env = 0.f;
      for (int t = t_hi; t >= 0; --t) {
        const long long n = q - static_cast<long long>(t) * kHop;
        if (n >= kNfft) break;
        const float w = win[n];
        acc += frames[(static_cast<size_t>(b) * Tmax + t) * 512 + n];
        env = fmaf(w, w, env);
      }
      v = env > 1e-11f ? acc / env * pk : 0.f;
    }
    out[b * out_stride + s] = v;
  }
}

}  // namespace

size_t stft_basis_floats() { return static_cast<size_t>(kNfft) * 512 + 512 * 512 + 512; }

void launch_stft_basis(float* basis, int sqrt_window, cudaStream_t s) {
  float* fwd = basis; float* inv = fwd + static_cast<size_t>(kNfft) * 512; float* win = inv + 512 * 512;
  launch_k(stft_basis_kernel, dim3((512 * 512 + 255) / 256), dim3(256), 0, s, fwd, inv, win, sqrt_window);
}

int stft_frames(int L) { return 1 + L / kHop; }

void launch_stft_spec(const float* basis, const float* wav, long long wav_stride, const int* lengths_dev, int B, int Lmax,
                      bool normalize, float factor, float expo, int mode, float* xpad, long long xpad_stride, float2* S,
                      unsigned* peak_bits, float2* Y, int Tpad, cudaStream_t s) {
  const int Tmax = stft_frames(Lmax);
  if (normalize) cudaMemsetAsync(peak_bits, 0, sizeof(unsigned) * B, s);
  launch_k(wav_prep_kernel, dim3(static_cast<unsigned>(std::min<long long>((xpad_stride + 255) / 256, 592)), B), dim3(256), 0, s,
           wav, wav_stride, lengths_dev, xpad, xpad_stride, normalize ? peak_bits : static_cast<unsigned*>(nullptr));
  SgemmArgs g{};
  g.A = xpad; g.lda = kHop; g.strideA = xpad_stride;            // frame t = xpad[t*128 .. t*128+510): no frame matrix
  g.Bm = basis; g.ldb = 512; g.strideB = 0; g.transB = 0;
  g.C = reinterpret_cast<float*>(S); g.ldc = 512; g.strideC = static_cast<long long>(Tmax) * 512;
  g.M = Tmax; g.N = 512; g.K = kNfft; g.batch = B; g.alpha = 1.f;
  launch_sgemm(g, s);
  launch_k(spec_fwd_kernel, dim3((Tpad + 31) / 32, kBins / 32, B), dim3(256), 0, s, static_cast<const float2*>(S), Tmax,
           lengths_dev, normalize ? static_cast<const unsigned*>(peak_bits) : static_cast<const unsigned*>(nullptr), factor,
           expo, mode, Y, Tpad);
}

void launch_spec_istft(const float* basis, const float2* X, int Tpad, const int* lengths_dev, int B, int Lmax, float factor,
                       float expo, int mode, const float* peak, float2* S, float* frames, float* wav_out, long long wav_stride,
                       cudaStream_t s) {
  (void)Lmax;
  const int Tmax = Tpad;                              // every frame of the padded spectrogram is synthesised
  const float* inv = basis + static_cast<size_t>(kNfft) * 512;
  const float* win = inv + 512 * 512;
  launch_k(spec_back_kernel, dim3((Tmax + 31) / 32, kBins / 32, B), dim3(256), 0, s, X, Tpad, factor, expo, mode, S, Tmax);
  SgemmArgs g{};
  g.A = reinterpret_cast<const float*>(S); g.lda = 512; g.strideA = static_cast<long long>(Tmax) * 512;
  g.Bm = inv; g.ldb = 512; g.strideB = 0; g.transB = 0;
  g.C = frames; g.ldc = 512; g.strideC = static_cast<long long>(Tmax) * 512;
  g.M = Tmax; g.N = 512; g.K = 512; g.batch = B; g.alpha = 1.f;
  launch_sgemm(g, s);
  launch_k(ola_kernel, dim3(static_cast<unsigned>(std::min<long long>((wav_stride + 255) / 256, 592)), B), dim3(256), 0, s,
           static_cast<const float*>(frames), Tmax, lengths_dev, win, peak, wav_out, wav_stride);
}

}  // namespace flowse
