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
};

int snn_launch_gray6(snn_ctx* ctx, const ExtractPlan& pl, int nf);
bool snn_gauss_fast_supported(int radius);
int snn_launch_gauss_fast(snn_ctx* ctx, const ExtractPlan& pl, int npass, int ppf);
int snn_launch_sobel_nms2(snn_ctx* ctx, const ExtractPlan& pl, int npass, int ppf, float* mag_dbg);
int snn_image_chain(snn_ctx* ctx, SnnChrom* c, const ExtractPlan& pl, int frame0, int nf, float* mag_dbg);
