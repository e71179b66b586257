// Image chain of getStripe.StripeSearch for batches of search windows:
//   k_frame_prepare  window fetch + NaN->0 + zero-column compaction     (getStripe.py:794-821)
//   k_gray           saturate -> brightness -> bf x bf mean (REFLECT_101) -> clip -> float32
//                    -> RGB2GRAY                                           (getStripe.py:889-913)
//   k_bleed          gaussian(ones) + eps, once per window (depends on S and sigma only)
//   k_gauss_v/_h     scipy.ndimage.gaussian_filter(mode='constant'): float32 storage, float64
//                    accumulation per axis, farthest pair first; then / bleed
//   k_sobel_nms      ndi.sobel x2 (mode reflect), float32 magnitude, bilinear non-maximum
//                    suppression, low/high thresholds -> two bit-packed masks
// Every floating-point operation is issued in the reference's order with explicit rounding
// (this TU is compiled with -fmad=false; the one FMA that OpenCV's gray conversion performs is
// written as __fmaf_rn).  Rooflines: k_frame_prepare is HBM-bound (band read once per
// window: 8 B per window pixel); k_gray..k_sobel_nms are FP64-pipe / L1-bound (≈ 90 fp64
// operations per pixel-pass), their HBM traffic is 4-12 B per pixel-pass of intermediates.
#include "image.cuh"

// ---- window geometry (getStripe.py:794-801) ------------------------------------------------
__host__ __device__ inline void frame_geometry(int idx, int64_t nbins, int* start, int* end) {
    int64_t s = (int64_t)idx * 200 - 100;
    int64_t e = ((int64_t)idx + 1) * 200 + 99;
    if (e >= nbins) e = nbins - 1;
    if (idx == 0) s = 0;
    *start = (int)s;
    *end = (int)e;
}

// The window fetch runs as three small kernels so that a batch of ~51 windows fills the GPU
// (one CTA per window left 2/3 of the SMs idle for 150 us per batch, profiles/r01b):
//   k_frame_flags    (window, 50-row slab): which columns hold a non-zero, non-NaN pixel
//   k_frame_compact  (window): ordered list of those columns, S, window geometry
//   k_frame_gather   (window, 8-row slab): Dc[a][b] = nan0(M[nz[a]][nz[b]])
#define FF_SLAB 50
__global__ void __launch_bounds__(256) k_frame_flags(const double* __restrict__ band, int64_t nbins, int bw,
                                                     int frame0, int* __restrict__ flag_all) {
    __shared__ int flag[SNN_SMAX];
    const int f = blockIdx.x;
    int start, end;
    frame_geometry(frame0 + f, nbins, &start, &end);
    const int n = end - start + 1;
    const int i0 = blockIdx.y * FF_SLAB;
    if (i0 >= n) return;
    const int i1 = min(i0 + FF_SLAB, n);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    for (int t = threadIdx.x; t < SNN_SMAX; t += blockDim.x) flag[t] = 0;
    __syncthreads();
    // colsum != 0  <=>  the column holds a non-zero, non-NaN pixel (values are non-negative)
    for (int i = i0 + warp; i < i1; i += nwarp) {
        const double* row = band + (int64_t)(start + i) * bw;
        bool any = false;
        for (int d = lane; d < n - i; d += 32) {
            const double v = row[d];
            if (v == v && v != 0.0) {
                any = true;
                flag[i + d] = 1;
            }
        }
        if (__any_sync(0xffffffffu, any) && lane == 0) flag[i] = 1;
    }
    __syncthreads();
    int* out = flag_all + (int64_t)f * SNN_SMAX;
    for (int t = threadIdx.x; t < n; t += blockDim.x)
        if (flag[t]) out[t] = 1;
}

__global__ void __launch_bounds__(32) k_frame_compact(const int* __restrict__ flag_all, int64_t nbins, int frame0,
                                                      FrameInfo* __restrict__ info, int32_t* __restrict__ nz_all) {
    const int f = blockIdx.x, lane = threadIdx.x;
    int start, end;
    frame_geometry(frame0 + f, nbins, &start, &end);
    const int n = end - start + 1;
    const int* flag = flag_all + (int64_t)f * SNN_SMAX;
    int32_t* nz = nz_all + (int64_t)f * SNN_SMAX;
    int S = 0;
    for (int t0 = 0; t0 < n; t0 += 32) {
        const int t = t0 + lane;
        const bool on = t < n && flag[t] != 0;
        const unsigned b = __ballot_sync(0xffffffffu, on);
        if (on) nz[S + __popc(b & ((1u << lane) - 1u))] = t;
        S += __popc(b);
    }
    const int Sraw = S;
    if (S <= 10) S = 0;  // getStripe.py:818 `if framesize > 10`
    __syncwarp();
    for (int t = S + lane; t < SNN_SMAX; t += 32) nz[t] = -1;
    (void)Sraw;
    if (lane == 0) {
        info[f].S = S;
        info[f].start = start;
        info[f].end = end;
        info[f].idx = frame0 + f;
    }
}

#define FG_ROWS 8
__global__ void __launch_bounds__(FG_ROWS * 32) k_frame_gather(const double* __restrict__ band, int bw,
                                                               const FrameInfo* __restrict__ info,
                                                               const int32_t* __restrict__ nz_all,
                                                               double* __restrict__ Dc_all) {
    __shared__ int nz[SNN_SMAX];
    const int f = blockIdx.x;
    const int S = info[f].S;
    const int a = blockIdx.y * FG_ROWS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (blockIdx.y * FG_ROWS >= S) return;
    for (int t = threadIdx.x; t < S; t += blockDim.x) nz[t] = nz_all[(int64_t)f * SNN_SMAX + t];
    __syncthreads();
    if (a >= S) return;
    const int start = info[f].start;
    const int ga = nz[a];
    const double* rowa = band + (int64_t)(start + ga) * (bw - 1) + start;   // M[ga][g] for g >= ga
    double* out = Dc_all + (int64_t)f * SNN_SMAX * SNN_SMAX + (int64_t)a * SNN_SMAX;
    for (int b = lane; b < S; b += 32) {
        const int gb = nz[b];
        // symmetric matrix element from the band: band[min * (bw - 1) + max] (see score_warp.cu)
        const double v = gb >= ga ? rowa[gb] : band[(int64_t)(start + gb) * (bw - 1) + start + ga];
        out[b] = snn_nan0(v);
    }
}

// ---- saturate + brightness + mean blur + gray ---------------------------------------------------
__device__ __forceinline__ int reflect101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return i;
}

__global__ void __launch_bounds__(256) k_gray(const double* __restrict__ Dc_all, const FrameInfo* __restrict__ info,
                                              GrayParams P, float* __restrict__ gray_all) {
    extern __shared__ double adj[];
    const int pass = blockIdx.z;
    const int f = pass / (P.nq * SNN_NBRIGHT);
    const int q = (pass / SNN_NBRIGHT) % P.nq;
    const int bi = pass % SNN_NBRIGHT;
    const int S = info[f].S;
    const int ty0 = blockIdx.y * 32, tx0 = blockIdx.x * 32;
    if (ty0 >= S || tx0 >= S) return;
    const double M = P.mp[q];
    const double b = P.bright[bi];
    const double k = (1.0 - 0.0) / (b - 0.0);
    const int bf = P.bf, a = bf / 2, T = 32 + bf - 1;
    const double* Dc = Dc_all + (int64_t)f * SNN_SMAX * SNN_SMAX;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    for (int t = tid; t < T * T; t += 256) {
        int yy = t / T, xx = t - yy * T;
        int r = ty0 + yy - a, c = tx0 + xx - a;
        double v = 0.0;
        if (r < S + bf && c < S + bf) {
            r = reflect101(r, S);
            c = reflect101(c, S);
            if (r >= 0 && r < S && c >= 0 && c < S) v = snn_adjust_pixel(Dc[r * SNN_SMAX + c], M, b, k);
        }
        adj[t] = v;
    }
    __syncthreads();
    float* gray = gray_all + (int64_t)pass * SNN_SMAX * SNN_SMAX;
    const float CR = 0.299f, CG = 0.587f, CB = 0.114f;
    const int x = threadIdx.x;
    if (tx0 + x >= S) return;
    for (int y = threadIdx.y; y < 32; y += 8) {
        if (ty0 + y >= S) break;
        double acc = 0.0;
        for (int ky = 0; ky < bf; ++ky)
            for (int kx = 0; kx < bf; ++kx) acc = acc + P.kv * adj[(y + ky) * T + (x + kx)];
        acc = fmin(fmax(acc, 0.0), 1.0);
        float g = (float)acc;
        float out;
        if (P.variant == SNN_GRAY_IPP) {
            float t = (P.r32 * CR) + (g * CG);
            out = __fmaf_rn(g, CB, t);
        } else if (P.variant == SNN_GRAY_FMA2) {
            float t = __fmaf_rn(g, CG, P.r32 * CR);
            out = __fmaf_rn(g, CB, t);
        } else {
            out = ((P.r32 * CR) + (g * CG)) + (g * CB);
        }
        gray[(ty0 + y) * SNN_SMAX + tx0 + x] = out;
    }
}

// ---- Gaussian ----------------------------------------------------------------------------------
// One pixel of gaussian(ones, mode='constant') + eps: scipy's two correlate1d passes over a mask of
// ones (float32 between the passes, float64 accumulation, farthest pair first).
__device__ __forceinline__ float bleed_at(const GaussParams& G, int i, int j, int S) {
    const int r = G.radius;
    double t = 1.0 * G.w[r];
    for (int k = r; k >= 1; --k) {
        double lo = (i - k >= 0) ? 1.0 : 0.0, hi = (i + k < S) ? 1.0 : 0.0;
        t = t + (lo + hi) * G.w[r - k];
    }
    const double bv = (double)(float)t;
    double u = bv * G.w[r];
    for (int k = r; k >= 1; --k) {
        double lo = (j - k >= 0) ? bv : 0.0, hi = (j + k < S) ? bv : 0.0;
        u = u + (lo + hi) * G.w[r - k];
    }
    return (float)u + 1.1920928955078125e-07f;
}

// Pixels at least `radius` away from every border see all taps: the same operations on the same
// operands for each of them (92 % of a 400 x 400 window), so a thread evaluates that value once
// and the loops only run for border pixels.
__global__ void __launch_bounds__(256) k_bleed(const FrameInfo* __restrict__ info, GaussParams G,
                                               float* __restrict__ bleed_all) {
    const int f = blockIdx.z;
    const int S = info[f].S;
    const int j = blockIdx.x * 32 + threadIdx.x;
    if (j >= S) return;
    const int r = G.radius;
    const float interior = bleed_at(G, r, r, 2 * r + 1);
    const bool jin = j >= r && j < S - r;
    for (int i = blockIdx.y * 32 + threadIdx.y; i < blockIdx.y * 32 + 32 && i < S; i += 8) {
        const bool in = jin && i >= r && i < S - r;
        bleed_all[(int64_t)f * SNN_SMAX * SNN_SMAX + i * SNN_SMAX + j] = in ? interior : bleed_at(G, i, j, S);
    }
}

// axis 0 (down the rows): out[i][j] = f32( x[i][j] w[r] + sum_{k=r..1} (x[i-k][j] + x[i+k][j]) w[r-k] )
__global__ void __launch_bounds__(256) k_gauss_v(const float* __restrict__ in_all, const FrameInfo* __restrict__ info,
                                                 int passes_per_frame, GaussParams G, float* __restrict__ out_all) {
    const int pass = blockIdx.z;
    const int S = info[pass / passes_per_frame].S;
    const int j = blockIdx.x * 32 + threadIdx.x;
    if (j >= S || blockIdx.y * 32 >= S) return;
    const float* in = in_all + (int64_t)pass * SNN_SMAX * SNN_SMAX;
    float* out = out_all + (int64_t)pass * SNN_SMAX * SNN_SMAX;
    const int r = G.radius;
    for (int i = blockIdx.y * 32 + threadIdx.y; i < blockIdx.y * 32 + 32 && i < S; i += 8) {
        double t = (double)in[i * SNN_SMAX + j] * G.w[r];
        for (int k = r; k >= 1; --k) {
            double lo = (i - k >= 0) ? (double)in[(i - k) * SNN_SMAX + j] : 0.0;
            double hi = (i + k < S) ? (double)in[(i + k) * SNN_SMAX + j] : 0.0;
            t = t + (lo + hi) * G.w[r - k];
        }
        out[i * SNN_SMAX + j] = (float)t;
    }
}

// axis 1 (along the row), then the bleed-over normalisation (float32 division)
__global__ void __launch_bounds__(256) k_gauss_h(const float* __restrict__ in_all, const FrameInfo* __restrict__ info,
                                                 int passes_per_frame, GaussParams G,
                                                 const float* __restrict__ bleed_all, float* __restrict__ out_all) {
    __shared__ float rowbuf[8][32 + 2 * 32];
    const int pass = blockIdx.z;
    const int f = pass / passes_per_frame;
    const int S = info[f].S;
    const int j0 = blockIdx.x * 32;
    if (j0 >= S || blockIdx.y * 32 >= S) return;
    const float* in = in_all + (int64_t)pass * SNN_SMAX * SNN_SMAX;
    const float* bleed = bleed_all + (int64_t)f * SNN_SMAX * SNN_SMAX;
    float* out = out_all + (int64_t)pass * SNN_SMAX * SNN_SMAX;
    const int r = G.radius;
    const int j = j0 + threadIdx.x;
    for (int ib = blockIdx.y * 32; ib < blockIdx.y * 32 + 32; ib += 8) {
        const int i = ib + threadIdx.y;
        __syncthreads();
        if (i < S) {
            for (int t = threadIdx.x; t < 32 + 2 * r; t += 32) {
                int c = j0 - r + t;
                rowbuf[threadIdx.y][t] = (c >= 0 && c < S) ? in[i * SNN_SMAX + c] : 0.f;
            }
        }
        __syncthreads();
        if (i < S && j < S) {
            const float* p = &rowbuf[threadIdx.y][threadIdx.x + r];
            double t = (double)p[0] * G.w[r];
            for (int k = r; k >= 1; --k) t = t + ((double)p[-k] + (double)p[k]) * G.w[r - k];
            out[i * SNN_SMAX + j] = (float)t / bleed[i * SNN_SMAX + j];
        }
    }
}

// ---- Sobel + magnitude + bilinear NMS + thresholds -------------------------------------------------
__device__ __forceinline__ int symm(int i, int n) {  // scipy 'reflect' (d c b a | a b c d | d c b a)
    if (i < 0) i = -i - 1;
    if (i >= n) i = 2 * n - 1 - i;
    return i;
}

__global__ void __launch_bounds__(256) k_sobel_nms(const float* __restrict__ sm_all, const FrameInfo* __restrict__ info,
                                                   int passes_per_frame, uint32_t* __restrict__ low_all,
                                                   uint32_t* __restrict__ high_all, float* __restrict__ mag_dbg) {
    __shared__ float sm[36][37];   // smoothed, tile + halo 2 (symmetric-reflected indices)
    __shared__ float si[34][35];   // isobel, tile + halo 1
    __shared__ float sj[34][35];   // jsobel
    __shared__ float mg[34][35];   // magnitude
    const int pass = blockIdx.z;
    const int S = info[pass / passes_per_frame].S;
    const int ty0 = blockIdx.y * 32, tx0 = blockIdx.x * 32;
    if (ty0 >= S) return;
    uint32_t* low = low_all + (int64_t)pass * SNN_SMAX * SNN_WPR;
    uint32_t* high = high_all + (int64_t)pass * SNN_SMAX * SNN_WPR;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    if (tx0 >= S) {  // words right of the image: zero
        if (tid < 32 && ty0 + tid < S) {
            low[(ty0 + tid) * SNN_WPR + blockIdx.x] = 0u;
            high[(ty0 + tid) * SNN_WPR + blockIdx.x] = 0u;
        }
        return;
    }
    const float* smo = sm_all + (int64_t)pass * SNN_SMAX * SNN_SMAX;
    for (int t = tid; t < 36 * 36; t += 256) {
        int yy = t / 36, xx = t - yy * 36;
        int r = ty0 + yy - 2, c = tx0 + xx - 2;
        float v = 0.f;
        if (r < S + 2 && c < S + 2) {
            r = symm(r, S);
            c = symm(c, S);
            if (r >= 0 && r < S && c >= 0 && c < S) v = smo[r * SNN_SMAX + c];
        }
        sm[yy][xx] = v;
    }
    __syncthreads();
    // Image coordinates beyond the border are handled by the reflected loads above: the
    // derivative pass of ndi.sobel is computed on the whole array first, so its value at a
    // reflected position equals the derivative of the reflected smoothed image EXCEPT that the
    // derivative direction must be taken at the source position.  Evaluate the two 1-D passes
    // explicitly for positions clamped into the image instead.
    for (int t = tid; t < 34 * 34; t += 256) {
        int yy = t / 34, xx = t - yy * 34;
        int r = ty0 + yy - 1, c = tx0 + xx - 1;  // image position (may be -1 or S: unused halo)
        float is = 0.f, js = 0.f, m = 0.f;
        if (r >= 0 && r < S && c >= 0 && c < S) {
            // tile-local index of (r, c) in sm is (yy + 1, xx + 1)
            const int Y = yy + 1, X = xx + 1;
            // isobel: derivative along axis 0, then [1 2 1] along axis 1 (columns c-1, c, c+1
            // reflected INTO the image: the derivative image is what gets padded)
            int Xl = (c - 1 < 0) ? X : X - 1, Xr = (c + 1 >= S) ? X : X + 1;
            int Yu = (r - 1 < 0) ? Y : Y - 1, Yd = (r + 1 >= S) ? Y : Y + 1;
            {
                float d0 = (float)(((double)sm[Y - 1][X] - (double)sm[Y + 1][X]) * -1.0);
                float dl = (float)(((double)sm[Y - 1][Xl] - (double)sm[Y + 1][Xl]) * -1.0);
                float dr = (float)(((double)sm[Y - 1][Xr] - (double)sm[Y + 1][Xr]) * -1.0);
                is = (float)((double)d0 * 2.0 + ((double)dl + (double)dr) * 1.0);
            }
            {
                float d0 = (float)(((double)sm[Y][X - 1] - (double)sm[Y][X + 1]) * -1.0);
                float du = (float)(((double)sm[Yu][X - 1] - (double)sm[Yu][X + 1]) * -1.0);
                float dd = (float)(((double)sm[Yd][X - 1] - (double)sm[Yd][X + 1]) * -1.0);
                js = (float)((double)d0 * 2.0 + ((double)du + (double)dd) * 1.0);
            }
            float mm = is * is;
            mm = mm + js * js;
            m = sqrtf(mm);
        }
        si[yy][xx] = is;
        sj[yy][xx] = js;
        mg[yy][xx] = m;
    }
    __syncthreads();
    const int x = threadIdx.x;
    for (int y = threadIdx.y; y < 32; y += 8) {
        const int r = ty0 + y, c = tx0 + x;
        bool keep = false, strong = false;
        if (r < S && c < S) {
            const int Y = y + 1, X = x + 1;
            const float m = mg[Y][X];
            if (mag_dbg) mag_dbg[(int64_t)pass * SNN_SMAX * SNN_SMAX + r * SNN_SMAX + c] = m;
            if (r >= 1 && r <= S - 2 && c >= 1 && c <= S - 2 && m >= 0.1f) {
                const float is = si[Y][X], js = sj[Y][X];
                const bool up = is >= 0.f, down = is <= 0.f, right = js >= 0.f, left = js <= 0.f;
                const bool cond1 = (up && right) || (down && left);
                const bool cond2 = (down && right) || (up && left);
                const float ai = fabsf(is), aj = fabsf(js);
                float w = 0.f, n11 = 0.f, n12 = 0.f, n21 = 0.f, n22 = 0.f;
                bool ok = true;
                if (cond1) {
                    if (ai > aj) {
                        w = aj / ai;
                        n11 = mg[Y + 1][X]; n12 = mg[Y + 1][X + 1]; n21 = mg[Y - 1][X]; n22 = mg[Y - 1][X - 1];
                    } else {
                        w = ai / aj;
                        n11 = mg[Y][X + 1]; n12 = mg[Y + 1][X + 1]; n21 = mg[Y][X - 1]; n22 = mg[Y - 1][X - 1];
                    }
                } else if (cond2) {
                    if (ai < aj) {
                        w = ai / aj;
                        n11 = mg[Y][X + 1]; n12 = mg[Y - 1][X + 1]; n21 = mg[Y][X - 1]; n22 = mg[Y + 1][X - 1];
                    } else {
                        w = aj / ai;
                        n11 = mg[Y - 1][X]; n12 = mg[Y - 1][X + 1]; n21 = mg[Y + 1][X]; n22 = mg[Y + 1][X - 1];
                    }
                } else {
                    ok = false;
                }
                if (ok) {
                    const double omw = 1.0 - (double)w;
                    const double plus = (double)(n12 * w) + (double)n11 * omw;
                    const double minus = (double)(n22 * w) + (double)n21 * omw;
                    const double md = (double)m;
                    keep = (plus <= md) && (minus <= md);
                    strong = keep && (m >= 0.2f);
                }
            }
        }
        unsigned lw = __ballot_sync(0xffffffffu, keep);
        unsigned hw = __ballot_sync(0xffffffffu, strong);
        if (x == 0 && r < S) {
            low[r * SNN_WPR + blockIdx.x] = lw;
            high[r * SNN_WPR + blockIdx.x] = hw;
        }
    }
}

// ---- host-side launch helpers ------------------------------------------------------------------------
int snn_image_chain(snn_ctx* ctx, SnnChrom* c, const ExtractPlan& pl, int frame0, int nf, float* mag_dbg) {
    const int npass = nf * pl.nq * SNN_NBRIGHT;
    const int ppf = pl.nq * SNN_NBRIGHT;
    {
        SnnScope scope(ctx, ST_FRAME, 3);
        SNN_CUDA(cudaMemsetAsync(pl.flags, 0, (size_t)nf * SNN_SMAX * sizeof(int), ctx->stream));
        k_frame_flags<<<dim3(nf, (SNN_SMAX + FF_SLAB - 1) / FF_SLAB), 256, 0, ctx->stream>>>(c->band, c->nbins, c->bw,
                                                                                           frame0, pl.flags);
        SNN_LAUNCH_CHECK("k_frame_flags");
        k_frame_compact<<<nf, 32, 0, ctx->stream>>>(pl.flags, c->nbins, frame0, pl.info, pl.nz);
        SNN_LAUNCH_CHECK("k_frame_compact");
        k_frame_gather<<<dim3(nf, (SNN_SMAX + FG_ROWS - 1) / FG_ROWS), FG_ROWS * 32, 0, ctx->stream>>>(
            c->band, c->bw, pl.info, pl.nz, pl.Dc);
        SNN_LAUNCH_CHECK("k_frame_gather");
    }
    const bool generic = ctx->force_generic;
    const int fused = generic ? 0 : snn_fused_mode(pl);
    if (fused == 2)          // two fused kernels, no bleed-over image (image_fused.cu)
        return snn_launch_fused(ctx, pl, npass, ppf, mag_dbg != nullptr, pl.bufA, pl.bufC, mag_dbg, true, true);
    {
        SnnScope scope(ctx, ST_FRAME, 1);
        k_bleed<<<dim3(13, 13, nf), dim3(32, 8), 0, ctx->stream>>>(pl.info, pl.gauss, pl.bleed);
        SNN_LAUNCH_CHECK("k_bleed");
    }
    if (fused == 1) {        // gray + axis-0 Gaussian fused; axis 1 and Sobel/NMS as separate kernels
        int rc = snn_launch_fused(ctx, pl, npass, ppf, mag_dbg != nullptr, pl.bufA, nullptr, nullptr, true, false);
        if (rc) return rc;
        {
            SnnScope scope(ctx, ST_GAUSS, 1);
            rc = snn_launch_gauss_fast(ctx, pl, npass, ppf, true);
            if (rc) return rc;
        }
        SnnScope scope(ctx, ST_NMS, 1);
        return snn_launch_sobel_nms2(ctx, pl, npass, ppf, mag_dbg);
    }
    {
        SnnScope scope(ctx, ST_GRAY, 1);
        if (!generic && pl.gray.bf == 3) {
            int rc = snn_launch_gray6(ctx, pl, nf);
            if (rc) return rc;
        } else {
            int T = 32 + pl.gray.bf - 1;
            k_gray<<<dim3(13, 13, npass), dim3(32, 8), (size_t)T * T * sizeof(double), ctx->stream>>>(
                pl.Dc, pl.info, pl.gray, pl.bufA);
            SNN_LAUNCH_CHECK("k_gray");
        }
    }
    {
        SnnScope scope(ctx, ST_GAUSS, 2);
        if (!generic && snn_gauss_fast_supported(pl.gauss.radius)) {
            int rc = snn_launch_gauss_fast(ctx, pl, npass, ppf);
            if (rc) return rc;
        } else {
            k_gauss_v<<<dim3(13, 13, npass), dim3(32, 8), 0, ctx->stream>>>(pl.bufA, pl.info, ppf, pl.gauss, pl.bufB);
            SNN_LAUNCH_CHECK("k_gauss_v");
            k_gauss_h<<<dim3(13, 13, npass), dim3(32, 8), 0, ctx->stream>>>(pl.bufB, pl.info, ppf, pl.gauss,
                                                                             pl.bleed, pl.bufC);
            SNN_LAUNCH_CHECK("k_gauss_h");
        }
    }
    {
        SnnScope scope(ctx, ST_NMS, 1);
        if (!generic) {
            int rc = snn_launch_sobel_nms2(ctx, pl, npass, ppf, mag_dbg);
            if (rc) return rc;
        } else {
            k_sobel_nms<<<dim3(13, 13, npass), dim3(32, 8), 0, ctx->stream>>>(pl.bufC, pl.info, ppf, pl.low, pl.high,
                                                                               mag_dbg);
            SNN_LAUNCH_CHECK("k_sobel_nms");
        }
    }
    return SNN_OK;
}
