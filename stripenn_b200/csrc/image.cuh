// Shared declarations of the window pipeline (image.cu, search.cu).
#pragma once
#include "snn_internal.cuh"

struct FrameInfo {
    int S;      // compacted window size, 0 = skipped (<= 10 non-empty columns)
    int start;  // first bin of the window
    int end;    // last bin of the window
    int idx;    // window id
};

struct GrayParams {
    int nq;
    int bf;
    int variant;
    float r32;      // the constant red channel after blur + clip + float32
    double kv;      // (np.ones((bf,bf)) / (bf*bf))[0,0]
    double bright[SNN_NBRIGHT];
    double kb[SNN_NBRIGHT];   // (1.0 - 0.0) / (bright - 0.0)  (ImageProcessing.py:15)
    double mp[64];  // maxpixel values
};

struct GaussParams {
    int radius;
    double w[33];  // w[0..radius]: left half of the symmetric kernel, w[radius] = centre
};

// Raw candidate as written by k_search (search.cu); finished by k_finish.
struct RawCand {
    int pass, ud, k;
    int x, y, right, hi;  // columns x..right, rows y..hi (window-local)
};

struct ExtractPlan {
    int nq;
    GrayParams gray;
    GaussParams gauss;
    int minH, maxW;
    // device buffers (per batch of `fb` windows)
    int fb;
    FrameInfo* info;
    int32_t* nz;      // fb x 400
    int* flags;       // fb x 400 non-empty column flags
    double* Dc;       // fb x 400 x 400
    float* bleed;     // fb x 400 x 400
    float* bufA;      // npass x 400 x 400 (gray)
    float* bufB;      // npass x 400 x 400 (after the axis-0 pass)
    float* bufC;      // npass x 400 x 400 (smoothed)
    uint32_t* low;    // npass x 400 x 13
    uint32_t* high;   // npass x 400 x 13
    int* counter;     // [0] = number of raw candidates, [1] = overflow flag
    RawCand* raw;
    int raw_cap;      // capacity of raw / cand (candidates of one snn_extract call)
    int batch_cap;    // upper bound of the candidates one batch can add
    snn_candidate* cand;  // finished candidates (device)
    // fused path (image_fused.cu): TMA tensor map of bufB (a CUtensorMap, kept opaque here)
    alignas(64) unsigned char tmap[128];
    bool tmap_ok;
};

int snn_launch_gray6(snn_ctx* ctx, const ExtractPlan& pl, int nf);
bool snn_gauss_fast_supported(int radius);
int snn_launch_gauss_fast(snn_ctx* ctx, const ExtractPlan& pl, int npass, int ppf, bool axis1_only = false);
int snn_launch_sobel_nms2(snn_ctx* ctx, const ExtractPlan& pl, int npass, int ppf, float* mag_dbg);
int snn_image_chain(snn_ctx* ctx, SnnChrom* c, const ExtractPlan& pl, int frame0, int nf, float* mag_dbg);

// ---- fused image chain (image_fused.cu)
int snn_fused_mode(const ExtractPlan& pl);
int snn_fused_make_tmap(snn_ctx* ctx, ExtractPlan* pl, size_t npass);
int snn_launch_fused(snn_ctx* ctx, const ExtractPlan& pl, int npass, int ppf, bool dbg, float* gray_dbg, float* sm_dbg,
                     float* mag_dbg, bool part_a, bool part_b);

#ifdef __CUDACC__
// One pixel of the saturated, brightness-adjusted image, the reference's literal operation order
// (getStripe.py:889-895, ImageProcessing.py:17-23)
__device__ __forceinline__ double snn_adjust_pixel(double D, double M, double b, double k) {
    double blue = (255.0 * (M - D)) / M;          // getStripe.py:889  255*(M - submat)/M
    if (blue < 0.0) blue = 0.0;                   // :890
    double g = blue / 255.0;                      // :894  blue / 255
    g = fmin(fmax(g, 0.0), 1.0);                  // :895  np.clip (NaN passes through)
    if (g != g) return 0.0;                       // imBrightness3D leaves NaN cells at 0
    if (g <= 0.0) return 0.0;                     // ImageProcessing.py:17-23
    if (g > b) return 1.0;
    return k * g;
}
#endif
