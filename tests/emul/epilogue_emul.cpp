// TEST INFRASTRUCTURE.  Host build of the per-element bodies in ava-256_b200/csrc/epilogue_body.h: the same inline
// functions the CUDA kernels call, driven by plain loops in place of the grid, so index arithmetic and rounding order
// can be checked against PyTorch on a machine without a GPU.  Never loaded by the product.
// Build: g++ -O1 -ffp-contract=off -shared -fPIC -I/usr/local/cuda/include -I<repo>/ava-256_b200/csrc (tests/emul/build.py)
#include "epilogue_body.h"

using namespace mvp_epi;

extern "C" {

int emul_composite_forward(int N, int H, int W, int V, const float *rayrgba, const float *ccw, const float *ccb, const float *bg,
                           float *irgbrec, float *rayalpha) {
    CompositeFwd a;
    a.HW = (size_t)H * W; a.rayrgba = rayrgba; a.ccw = ccw; a.ccb = ccb; a.bg = bg; a.irgbrec = irgbrec; a.rayalpha = rayalpha;
    if (V == 4 && a.HW % 4) return -1;
    for (int n = 0; n < N; ++n)
        for (size_t px = 0; px < a.HW; px += V) V == 4 ? composite_fwd<4>(a, n, px) : composite_fwd<1>(a, n, px);
    return 0;
}

int emul_composite_backward(int N, int H, int W, int V, const float *rayrgba, const float *ccw, const float *bg,
                            const float *grad_irgbrec, const float *grad_rayalpha, float *grad_rayrgba, float *grad_ccw,
                            float *grad_ccb, float *grad_bg) {
    CompositeBwd a;
    a.HW = (size_t)H * W; a.rayrgba = rayrgba; a.ccw = ccw; a.bg = bg; a.grad_irgbrec = grad_irgbrec;
    a.grad_rayalpha = grad_rayalpha; a.grad_rayrgba = grad_rayrgba; a.grad_bg = grad_bg; a.want_cc = grad_ccw != nullptr;
    if (V == 4 && a.HW % 4) return -1;
    for (int n = 0; n < N; ++n) {
        double acc[6] = {0, 0, 0, 0, 0, 0};
        for (size_t px = 0; px < a.HW; px += V) {
            float part[6] = {0, 0, 0, 0, 0, 0};
            V == 4 ? composite_bwd<4>(a, n, px, part) : composite_bwd<1>(a, n, px, part);
            for (int i = 0; i < 6; ++i) acc[i] += part[i];
        }
        if (grad_ccw)
            for (int i = 0; i < 3; ++i) { grad_ccw[n * 3 + i] += (float)acc[i]; grad_ccb[n * 3 + i] += (float)acc[3 + i]; }
    }
    return 0;
}

static size_t payload_total(int N, int hb, int wb, int B, int V) { return (size_t)N * hb * B * B * ((size_t)wb * B) / V; }

int emul_payload_forward(int N, int hb, int wb, int B, int V, int BT, const float *tex, const float *opacity, float rgb_scale,
                         float rgb_bias, float *tplate) {
    PayloadArgs a = {};
    a.hb = hb; a.wb = wb; a.B = B; a.rgb_scale = rgb_scale; a.rgb_bias = rgb_bias; a.tex = tex; a.opacity = opacity; a.tplate = tplate;
    if ((V == 4 && B % 4) || (BT && BT != B)) return -1;
    const size_t total = payload_total(N, hb, wb, B, V);
    for (size_t e = 0; e < total; ++e) {
        if (V == 4 && BT == 8) payload_fwd<4, 8>(a, e);
        else if (V == 4) payload_fwd<4, 0>(a, e);
        else payload_fwd<1, 0>(a, e);
    }
    return 0;
}

int emul_payload_backward(int N, int hb, int wb, int B, int V, int BT, const float *tplate, const float *grad_tplate,
                          float rgb_scale, float *grad_tex, float *grad_opacity) {
    PayloadArgs a = {};
    a.hb = hb; a.wb = wb; a.B = B; a.rgb_scale = rgb_scale; a.tplate_in = tplate; a.grad_tplate = grad_tplate;
    a.grad_tex = grad_tex; a.grad_opacity = grad_opacity;
    if ((V == 4 && B % 4) || (BT && BT != B)) return -1;
    const size_t total = payload_total(N, hb, wb, B, V);
    for (size_t e = 0; e < total; ++e) {
        if (V == 4 && BT == 8) payload_bwd<4, 8>(a, e);
        else if (V == 4) payload_bwd<4, 0>(a, e);
        else payload_bwd<1, 0>(a, e);
    }
    return 0;
}

}  // extern "C"
