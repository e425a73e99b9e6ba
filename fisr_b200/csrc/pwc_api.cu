// C ABI of the PWC-Net inference path (include/fisr_b200.h, "PWC-Net"): parameter store under the TensorFlow variable names of
// philferriere/tfoptflow (scope pwcnet/), one plan (buffers + launch list) per input size, forward of N image pairs.
// Reference: FISR_tfoptflow/model_pwcnet.py:1012-1593 driven by FISR_for_video_pwcnet_predict_from_img_test.py:96-139.
// Routing of the 3x3 convs (build_plan): everything with >= 16 outputs on an image of >= 4 x 4 pixels goes to the tcgen05 conv kernel in
// split mode through build_split_conv (conv_umma.cu) -- dense-block channel slices as TMA boxes that start at a channel offset, dilated
// layers as polyphase launches, stride-2 layers as two row-phase launches (build_stride2), the flow predictor fused with the next
// level's up_feat transposed conv (build_fused), the 16-channel level-1 layers on 4-pixel super-pixels (build_packed4); the rest runs on
// the CUDA-core kernels of pwc_kernels.cu.  Activations are fp16 (hi, lo) planes throughout.
#include <algorithm>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/fisr_b200.h"
#include "aux_kernels.h"
#include "conv_umma.h"
#include "pwc_kernels.h"

using namespace fisr::pwc;
using fisr::ConvLaunch;

namespace {

constexpr int kLvls = 6, kPredLvl = 2, kCorr = 81, kCorrPad = 88, kGap = kCorrPad - kCorr;   // search range 4; corr slot padded to a multiple of 8
constexpr int kTail = 8;                                                // up_flow (2) + up_feat (2) + 4 zero channels: 16-byte pixel rows for TMA
constexpr int kChann[7] = {0, 16, 32, 64, 96, 128, 196};                // model_pwcnet.py:1083
constexpr int kDense[5] = {128, 128, 96, 64, 32};                       // model_pwcnet.py:1415-1433
constexpr int kActs = 448;                                              // 128 + 128 + 96 + 64 + 32
constexpr int kCtxtF[7] = {128, 128, 128, 96, 64, 32, 2};               // model_pwcnet.py:1506-1519
constexpr int kCtxtD[7] = {1, 2, 4, 8, 16, 1, 1};

struct PDef {
    std::string name;        // without /kernel, /bias
    int cin, cout;           // reference channel counts
    bool transpose;          // conv2d_transpose kernel [4,4,out,in]
    int gap_at;              // reference input channel index after which kGap zero rows are inserted (-1: none)
};

// channels of the dense buffer D_l: [act4 | act3 | act2 | act1 | act0 | corr (81 + 7 pad) | c1 | up_flow | up_feat | 4 pad]
int dense_cs(int lvl) { return kActs + kCorrPad + (lvl == kLvls ? 0 : kChann[lvl] + kTail); }
int pad8(int c) { return (c + 7) / 8 * 8; }
int dense_off(int k) { int o = kActs; for (int i = 0; i <= k; ++i) o -= kDense[i]; return o; }      // where act_k is written
int dense_in_off(int k) { return k == 0 ? kActs : dense_off(k - 1); }                               // where conv_k starts reading

std::vector<PDef> build_inventory() {
    std::vector<PDef> v;
    for (int l = 1; l <= kLvls; ++l) {
        const std::string p = "pwcnet/featpyr/conv" + std::to_string(l);
        v.push_back({p + "a", l == 1 ? 3 : kChann[l - 1], kChann[l], false, -1});
        v.push_back({p + "aa", kChann[l], kChann[l], false, -1});
        v.push_back({p + "b", kChann[l], kChann[l], false, -1});
    }
    for (int l = kLvls; l >= kPredLvl; --l) {
        int c = kCorr + (l == kLvls ? 0 : kChann[l] + 4);
        for (int k = 0; k < 5; ++k) {
            v.push_back({"pwcnet/predict_flow/conv" + std::to_string(l) + "_" + std::to_string(k), c, kDense[k], false, c - (l == kLvls ? 0 : kChann[l] + 4)});
            c += kDense[k];
        }
        const int gap = c - (l == kLvls ? 0 : kChann[l] + 4);
        v.push_back({"pwcnet/predict_flow/flow" + std::to_string(l), c, 2, false, gap});
        int cc = c;
        for (int k = 0; k < 7; ++k) {
            v.push_back({"pwcnet/ctxt/dc_conv" + std::to_string(l) + std::to_string(k + 1), cc, kCtxtF[k], false, k == 0 ? gap : -1});
            cc = kCtxtF[k];
        }
        if (l != kPredLvl) {
            v.push_back({"pwcnet/upsample/up_flow" + std::to_string(l), 2, 2, true, -1});
            v.push_back({"pwcnet/upsample/up_feat" + std::to_string(l), c, 2, true, gap});
        }
    }
    return v;
}
const std::vector<PDef>& inventory() {
    static const std::vector<PDef> inv = build_inventory();
    return inv;
}
const std::vector<std::string>& names() {
    static std::vector<std::string> n;
    if (n.empty())
        for (const auto& d : inventory()) { n.push_back(d.name + "/kernel"); n.push_back(d.name + "/bias"); }
    return n;
}

struct Param {
    int cin_pad = 0;             // input channels as the kernels see them (reference count + kGap where a gap is inserted)
    int cout = 0;
    float* d_w = nullptr;        // kernel layout: conv [9][cin_pad][cout]; transpose [16][2][cin_pad]
    float* d_b = nullptr;
    // tensor-core path (stride-1 3x3 convs with >= 16 outputs): fp16 (hi, lo) operand planes and the bias padded to cout_pad
    int KB = 0, cout_pad = 0;
    __half* d_wp = nullptr;
    float* d_bp = nullptr;
    bool packed = false;
    std::vector<float> h_w, h_b; // as uploaded (reference layout)
};

struct Op {
    bool umma = false;
    int group = 0;                                     // > 0: launches with the same id are independent (the polyphase launches of one dilated layer)
    int lane = 0;                                      // 1: the feature pyramid of image 2, which runs beside image 1's on a side stream
    ConvLaunch conv;                                   // umma: one tcgen05 conv launch (tensor maps encoded at plan build)
    std::function<void(cudaStream_t)> fn;              // otherwise
};

struct Plan {
    int N = 0, H = 0, W = 0;
    std::vector<void*> allocs;
    std::vector<Op> ops;
    int umma_ops = 0, groups = 0;
    const float *img1 = nullptr, *img2 = nullptr;     // bound per call
    float* out = nullptr;
    float* flow[kLvls + 1] = {nullptr};
    ~Plan() { for (void* p : allocs) cudaFree(p); }
};

}  // namespace

struct fisr_pwc {
    int device = 0, num_sms = 0;
    cudaStream_t stream = nullptr;
    fisr::EncodeTiledFn encode = nullptr;
    // independent launches of one layer are spread over the caller's stream and these, so that the partial last wave of one
    // persistent conv kernel overlaps the first wave of the next (each launch fills the GPU for only 1.3 - 3.6 waves)
    static constexpr int kSide = 3;
    cudaStream_t side[kSide] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[kSide] = {nullptr, nullptr, nullptr};
    int* d_err = nullptr;        // barrier-timeout flag of the conv kernel
    int* h_err = nullptr;        // pinned copy, refreshed at the end of every forward
    int use_umma = 3;            // FISR_PWC_UMMA=0: every conv on the CUDA-core kernel; 1: tensor cores except the dilated layers; 2: + dilated
                                 // layers as polyphase launches; 3 (default): + those launches spread over 4 streams (A/B measurements)
    std::vector<Param> params;
    Param fused[kLvls + 1];      // predict_flow/flow<l> and upsample/up_feat<l> as ONE 3x3 conv with 16 output columns (build_fused)
    Param s2[kLvls + 1][2];      // featpyr/conv<l>a (stride 2) as two stride-1 convs on the row phases of the input (build_stride2)
    Param pk4[2];                // featpyr/conv1aa, conv1b (16 -> 16) on 4-pixel super-pixels: 64 -> 64 (build_packed4)
    std::map<std::string, int> index;
    std::map<std::string, std::unique_ptr<Plan>> plans;
    Plan* last = nullptr;
    long long launches = 0;
    std::string err;
};

namespace {

std::string g_err;
int fail(fisr_pwc* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_err = buf;
    return code;
}
#define PWC_TRY(c, expr)                                                                                                       \
    do {                                                                                                                        \
        cudaError_t e__ = (expr);                                                                                               \
        if (e__ != cudaSuccess) return fail(c, FISR_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

struct Guard {
    int prev = -1;
    explicit Guard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~Guard() { if (prev >= 0) cudaSetDevice(prev); }
};

struct View { Planes b; int coff, C; };      // channels [coff, coff + C) of a plane buffer

int ensure_packed(fisr_pwc* c, Param& p) {
    if (p.packed) return FISR_OK;
    p.KB = (p.cin_pad + 63) / 64;
    p.cout_pad = fisr::split_conv_cout_pad(p.cout);
    if (!p.d_wp) {
        PWC_TRY(c, cudaMalloc(&p.d_wp, static_cast<size_t>(2) * p.KB * 9 * p.cout_pad * 64 * sizeof(__half)));
        PWC_TRY(c, cudaMalloc(&p.d_bp, p.cout_pad * sizeof(float)));
    }
    PWC_TRY(c, cudaMemsetAsync(p.d_bp, 0, p.cout_pad * sizeof(float), c->stream));
    PWC_TRY(c, cudaMemcpyAsync(p.d_bp, p.d_b, p.cout * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
    fisr::launch_prep_weights(p.d_w, p.d_wp, p.cin_pad, p.cout, p.KB, p.cout_pad, 2, c->stream);
    PWC_TRY(c, cudaGetLastError());
    p.packed = true;
    return FISR_OK;
}

// The flow predictor (3x3, 2 outputs) and the up_feat transposed conv (4x4, stride 2, 2 outputs) both read the whole dense buffer
// of a level.  out[2i + a] of the transposed conv gathers in[i + dy] w[ky] with (a, dy) -> ky: (0, 0) -> 1, (0, -1) -> 3,
// (1, 0) -> 2, (1, +1) -> 0 (same in x), i.e. it is a 3x3 conv at input resolution with one output column per (sub-pixel, filter):
// both layers run as ONE tensor-core conv with 16 columns, [flow 0..1 | (2a + b) * 2 + co of up_feat | 6 zeros].
int build_fused(fisr_pwc* c, int l) {
    Param& f = c->fused[l];
    if (f.packed) return FISR_OK;
    const std::string sl = std::to_string(l);
    const Param& pf = c->params[c->index.at("pwcnet/predict_flow/flow" + sl)];
    const PDef& df = inventory()[c->index.at("pwcnet/predict_flow/flow" + sl)];
    f.cin_pad = pf.cin_pad;
    f.cout = 16;
    std::vector<float> w(static_cast<size_t>(9) * f.cin_pad * 16, 0.f), b(16, 0.f);
    auto slot = [&](int ci) { return (df.gap_at >= 0 && ci >= df.gap_at) ? ci + kGap : ci; };
    if (!pf.h_w.empty())
        for (int t = 0; t < 9; ++t)
            for (int ci = 0; ci < df.cin; ++ci)
                for (int co = 0; co < 2; ++co) w[(static_cast<size_t>(t) * f.cin_pad + slot(ci)) * 16 + co] = pf.h_w[(static_cast<size_t>(t) * df.cin + ci) * 2 + co];
    if (!pf.h_b.empty()) { b[0] = pf.h_b[0]; b[1] = pf.h_b[1]; }
    if (l != kPredLvl) {
        const Param& pe = c->params[c->index.at("pwcnet/upsample/up_feat" + sl)];
        const PDef& de = inventory()[c->index.at("pwcnet/upsample/up_feat" + sl)];
        static const int kTap[2][3] = {{3, 1, -1}, {-1, 2, 0}};          // [sub-pixel parity][dy + 1] -> transposed-conv tap, -1: none
        if (!pe.h_w.empty())
            for (int sa = 0; sa < 2; ++sa)
                for (int sb = 0; sb < 2; ++sb)
                    for (int dy = 0; dy < 3; ++dy)
                        for (int dx = 0; dx < 3; ++dx) {
                            const int ky = kTap[sa][dy], kx = kTap[sb][dx];
                            if (ky < 0 || kx < 0) continue;
                            for (int co = 0; co < 2; ++co)
                                for (int ci = 0; ci < de.cin; ++ci)           // reference layout [4,4,out,in]
                                    w[(static_cast<size_t>(dy * 3 + dx) * f.cin_pad + slot(ci)) * 16 + 2 + (2 * sa + sb) * 2 + co] =
                                        pe.h_w[(static_cast<size_t>(ky * 4 + kx) * 2 + co) * de.cin + ci];
                        }
        if (!pe.h_b.empty())
            for (int s4 = 0; s4 < 4; ++s4) { b[2 + 2 * s4] = pe.h_b[0]; b[3 + 2 * s4] = pe.h_b[1]; }
    }
    if (!f.d_w) {
        PWC_TRY(c, cudaMalloc(&f.d_w, w.size() * sizeof(float)));
        PWC_TRY(c, cudaMalloc(&f.d_b, 16 * sizeof(float)));
    }
    PWC_TRY(c, cudaMemcpy(f.d_w, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice));
    PWC_TRY(c, cudaMemcpy(f.d_b, b.data(), 16 * sizeof(float), cudaMemcpyHostToDevice));
    return ensure_packed(c, f);
}

// A stride-2 3x3 conv ('same' on an even size: taps x[2o + k], k = 0..2) on the tensor cores.  View the input as super-pixels:
// pixel X of row phase py holds the 2C channels of input pixels (2Y + py, 2X) and (2Y + py, 2X + 1), which are contiguous in a compact
// NHWC buffer.  Then  out[Y, X] = sum over py of a stride-1 3x3 conv on that view whose only non-zero taps are
//   rows:  py = 0: dy' = 0 (ky = 0), dy' = +1 (ky = 2);   py = 1: dy' = 0 (ky = 1)
//   cols:  channel half 0: dx' = 0 (kx = 0), dx' = +1 (kx = 2);   channel half 1: dx' = 0 (kx = 1)
// Two launches: phase 0 writes bias + partial sum to an fp32 scratch, phase 1 adds it as its residual and applies the activation.
int build_stride2(fisr_pwc* c, int l) {
    const std::string name = "pwcnet/featpyr/conv" + std::to_string(l) + "a";
    const Param& src = c->params[c->index.at(name)];
    const PDef& d = inventory()[c->index.at(name)];
    const int C = d.cin, co = d.cout;
    for (int py = 0; py < 2; ++py) {
        Param& f = c->s2[l][py];
        if (f.packed) continue;
        f.cin_pad = 2 * C;
        f.cout = co;
        std::vector<float> w(static_cast<size_t>(9) * 2 * C * co, 0.f), b(co, 0.f);
        if (!src.h_w.empty())
            for (int ky = py; ky < 3; ky += 2)                  // py = 0: ky 0, 2;  py = 1: ky 1
                for (int kx = 0; kx < 3; ++kx) {
                    const int ty = ky / 2 + 1, tx = kx / 2 + 1, half = kx & 1;
                    for (int ci = 0; ci < C; ++ci)
                        for (int o = 0; o < co; ++o)
                            w[(static_cast<size_t>(ty * 3 + tx) * 2 * C + half * C + ci) * co + o] = src.h_w[(static_cast<size_t>(ky * 3 + kx) * C + ci) * co + o];
                }
        if (py == 0 && !src.h_b.empty()) b = src.h_b;
        if (!f.d_w) {
            PWC_TRY(c, cudaMalloc(&f.d_w, w.size() * sizeof(float)));
            PWC_TRY(c, cudaMalloc(&f.d_b, co * sizeof(float)));
        }
        PWC_TRY(c, cudaMemcpy(f.d_w, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice));
        PWC_TRY(c, cudaMemcpy(f.d_b, b.data(), co * sizeof(float), cudaMemcpyHostToDevice));
        const int rc = ensure_packed(c, f);
        if (rc != FISR_OK) return rc;
    }
    return FISR_OK;
}

// A 16 -> 16 conv on a compact [H, W, 16] tensor read as [H, W/4, 64]: super-pixel X' holds pixels 4X' .. 4X' + 3.  Output sub-pixel q,
// tap kx reads input pixel 4X' + q + kx - 1 = sub-pixel (q + kx - 1) mod 4 of super-pixel X' + floor((q + kx - 1) / 4): a 3x3 conv with
// 64 input and 64 output channels whose weights are a block-sparse copy of the 16 x 16 ones.  4x the MMA work (trivial at K = 144), but
// the TMA boxes are all real data: on the 16-channel view a 64-channel box fetched 128 B per pixel for 32 B of tensor (ncu: 1.96 GB of
// L2 -> SM traffic for 267 MB of input) and the level-1 convs ran at the L2 rate.
int build_packed4(fisr_pwc* c, int which, const std::string& name) {
    Param& f = c->pk4[which];
    if (f.packed) return FISR_OK;
    const Param& src = c->params[c->index.at(name)];
    f.cin_pad = 64;
    f.cout = 64;
    std::vector<float> w(static_cast<size_t>(9) * 64 * 64, 0.f), b(64, 0.f);
    if (!src.h_w.empty())
        for (int ky = 0; ky < 3; ++ky)
            for (int kx = 0; kx < 3; ++kx)
                for (int q = 0; q < 4; ++q) {
                    const int s = q + kx - 1, dx = s < 0 ? -1 : (s > 3 ? 1 : 0), qi = s - 4 * dx;
                    for (int ci = 0; ci < 16; ++ci)
                        for (int co = 0; co < 16; ++co)
                            w[(static_cast<size_t>(ky * 3 + dx + 1) * 64 + qi * 16 + ci) * 64 + q * 16 + co] = src.h_w[(static_cast<size_t>(ky * 3 + kx) * 16 + ci) * 16 + co];
                }
    if (!src.h_b.empty())
        for (int q = 0; q < 4; ++q)
            for (int co = 0; co < 16; ++co) b[q * 16 + co] = src.h_b[co];
    if (!f.d_w) {
        PWC_TRY(c, cudaMalloc(&f.d_w, w.size() * sizeof(float)));
        PWC_TRY(c, cudaMalloc(&f.d_b, 64 * sizeof(float)));
    }
    PWC_TRY(c, cudaMemcpy(f.d_w, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice));
    PWC_TRY(c, cudaMemcpy(f.d_b, b.data(), 64 * sizeof(float), cudaMemcpyHostToDevice));
    return ensure_packed(c, f);
}

int build_plan(fisr_pwc* c, int N, int H, int W, Plan** out) {
    if (N < 1 || H < 64 || W < 64 || H % 64 || W % 64)
        return fail(c, FISR_E_INVALID, "PWC-Net (6-level pyramid) needs H, W multiples of 64 (got %d x %d x %d): pad like adapt_x", N, H, W);
    char key[48];
    snprintf(key, sizeof key, "%d_%d_%d", N, H, W);
    auto it = c->plans.find(key);
    if (it != c->plans.end()) { *out = it->second.get(); return FISR_OK; }
    std::unique_ptr<Plan> plan(new Plan());
    plan->N = N; plan->H = H; plan->W = W;
    Plan* pl = plan.get();
    int rc = FISR_OK;
    auto alloc_bytes = [&](size_t bytes) -> void* {
        void* p = nullptr;
        if (rc != FISR_OK) return nullptr;
        if (cudaMalloc(&p, bytes) != cudaSuccess) { rc = fail(c, FISR_E_NOMEM, "cudaMalloc(%zu bytes) failed", bytes); return nullptr; }
        cudaMemsetAsync(p, 0, bytes, c->stream);      // gap / pad channels stay zero for ever
        plan->allocs.push_back(p);
        return p;
    };
    auto px = [&](int l) { return static_cast<size_t>(N) * (H >> l) * (W >> l); };
    auto planes = [&](size_t npix, int cs) -> Planes {
        Planes b{};
        b.cs = cs;
        b.plane = (npix * cs + 511) / 512 * 512;
        b.p = static_cast<__half*>(alloc_bytes(2 * b.plane * sizeof(__half)));
        return b;
    };
    auto f32 = [&](size_t n) { return static_cast<float*>(alloc_bytes(n * sizeof(float))); };
    auto P = [&](const std::string& n) -> Param& { return c->params[c->index.at(n)]; };
    int lane = 0;
    float* s2tmp[2] = {nullptr, nullptr};      // fp32 partial sums of the two-launch stride-2 convs, one per pyramid lane
    auto push = [&](std::function<void(cudaStream_t)> fn) { Op op; op.fn = std::move(fn); op.lane = lane; pl->ops.push_back(std::move(op)); };
    // One conv layer: on the tensor cores when it is a stride-1 conv with >= 16 outputs on an image of at least 4 x 4 pixels --
    // dilation d as d (column phase) x d (row phase) undilated convs on the polyphase sub-images, each launch covering the d
    // column phases of one row phase of one image as a batch -- else on the CUDA-core kernel.
    auto conv_p = [&](Param& p, const std::string& name, View in, View outv, int l_in, int l_out, int stride, int dil, bool leaky, float* out_f32 = nullptr,
                      const float* add = nullptr) {
        if (rc != FISR_OK) return;
        const int Hin = H >> l_in, Win = W >> l_in, Hout = H >> l_out, Wout = W >> l_out;
        const bool divisible = Hout % dil == 0 && Wout % dil == 0;
        if (c->use_umma >= 1 && stride == 1 && dil == 1 && !out_f32 && p.cin_pad == 16 && p.cout == 16 && in.coff == 0 && in.b.cs == 16 && outv.coff == 0 &&
            outv.b.cs == 16 && Wout % 4 == 0 && Wout / 4 >= 4 && Hout >= 4 && (name == "pwcnet/featpyr/conv1aa" || name == "pwcnet/featpyr/conv1b")) {
            const int which = name == "pwcnet/featpyr/conv1b";
            if ((rc = build_packed4(c, which, name)) != FISR_OK) return;
            Param& q = c->pk4[which];
            const int wp = Wout / 4;
            fisr::SplitConvDesc d{};
            d.in = in.b.p; d.in_plane = in.b.plane; d.in_cs = 64;
            d.in_sx = 64; d.in_sy = static_cast<long long>(Wout) * 16; d.in_sn = static_cast<long long>(Hout) * Wout * 16;
            d.cin_off = 0; d.cin = 64;
            d.out = outv.b.p; d.out_plane = outv.b.plane; d.out_cs = 64; d.out_off = 0;
            d.opix_x = 1; d.opix_y = wp; d.opix_n = static_cast<long long>(Hout) * wp;
            d.N = N; d.H = Hout; d.W = wp;
            d.out_pixels = static_cast<long long>(N) * Hout * wp;
            d.wp = q.d_wp; d.bias = q.d_bp; d.cout = 64; d.cout_pad = q.cout_pad;
            d.relu = 0; d.slope = leaky ? 0.1f : 0.f;
            Op op;
            op.umma = true;
            op.lane = lane;
            std::string why;
            if (!fisr::build_split_conv(c->encode, d, c->num_sms, c->d_err, &op.conv, &why)) {
                rc = fail(c, FISR_E_CUDA, "conv %s (4-pixel packing): %s", name.c_str(), why.c_str());
                return;
            }
            pl->ops.push_back(std::move(op));
            pl->umma_ops++;
            return;
        }
        if (c->use_umma >= (dil == 1 ? 1 : 2) && stride == 1 && !out_f32 && p.cout >= 16 && divisible && Hout / dil >= 4 && Wout / dil >= 4) {
            if ((rc = ensure_packed(c, p)) != FISR_OK) return;
            const int hs = Hout / dil, ws = Wout / dil;
            const long long pixrow = static_cast<long long>(Wout), piximg = static_cast<long long>(Hout) * Wout;
            const int group = dil == 1 ? 0 : ++pl->groups;
            for (int n = 0; n < (dil == 1 ? 1 : N); ++n)
                for (int py = 0; py < dil; ++py) {
                    fisr::SplitConvDesc d{};
                    const long long pix0 = dil == 1 ? 0 : n * piximg + py * pixrow;          // first pixel of this launch's lattice
                    d.in = in.b.p + pix0 * in.b.cs; d.in_plane = in.b.plane; d.in_cs = in.b.cs;
                    d.cin_off = in.coff; d.cin = p.cin_pad;
                    d.out = outv.b.p + pix0 * outv.b.cs; d.out_plane = outv.b.plane; d.out_cs = outv.b.cs; d.out_off = outv.coff;
                    if (dil == 1) {
                        d.N = N; d.in_sx = in.b.cs; d.in_sy = pixrow * in.b.cs; d.in_sn = piximg * in.b.cs;
                        d.opix_x = 1; d.opix_y = pixrow; d.opix_n = piximg;
                    } else {          // image index = column phase
                        d.N = dil; d.in_sx = static_cast<long long>(dil) * in.b.cs; d.in_sy = dil * pixrow * in.b.cs; d.in_sn = in.b.cs;
                        d.opix_x = dil; d.opix_y = dil * pixrow; d.opix_n = 1;
                    }
                    d.H = hs; d.W = ws;
                    d.out_pixels = static_cast<long long>(N) * piximg;
                    d.wp = p.d_wp; d.bias = p.d_bp; d.cout = p.cout; d.cout_pad = p.cout_pad;
                    d.relu = 0; d.slope = leaky ? 0.1f : 0.f;
                    Op op;
                    op.umma = true;
                    op.group = group;
                    op.lane = lane;
                    std::string why;
                    if (!fisr::build_split_conv(c->encode, d, c->num_sms, c->d_err, &op.conv, &why)) {
                        rc = fail(c, FISR_E_CUDA, "conv %s (%d x %d, dilation %d): %s", name.c_str(), Hout, Wout, dil, why.c_str());
                        return;
                    }
                    pl->ops.push_back(std::move(op));
                    pl->umma_ops++;
                }
            return;
        }
        if (c->use_umma >= 1 && stride == 2 && l_in >= 1 && !out_f32 && Hout >= 4 && Wout >= 4 && in.coff == 0 && in.b.cs == in.C &&
            in.C % 8 == 0 && p.cout % 4 == 0 && p.cout > 16 && name.find("featpyr/conv") != std::string::npos) {
            const int l = l_out;
            if ((rc = build_stride2(c, l)) != FISR_OK) return;
            const long long pixrow = static_cast<long long>(Wout), piximg = static_cast<long long>(Hout) * Wout;
            float* tmp = s2tmp[lane];
            for (int py = 0; py < 2; ++py) {
                Param& q = c->s2[l][py];
                fisr::SplitConvDesc d{};
                d.in = in.b.p + static_cast<long long>(py) * Win * in.b.cs; d.in_plane = in.b.plane; d.in_cs = 2 * in.b.cs;
                d.in_sx = 2LL * in.b.cs; d.in_sy = 2LL * Win * in.b.cs; d.in_sn = static_cast<long long>(Hin) * Win * in.b.cs;
                d.cin_off = 0; d.cin = q.cin_pad;
                d.out = outv.b.p; d.out_plane = outv.b.plane; d.out_cs = outv.b.cs; d.out_off = outv.coff;
                d.opix_x = 1; d.opix_y = pixrow; d.opix_n = piximg;
                d.N = N; d.H = Hout; d.W = Wout;
                d.out_pixels = static_cast<long long>(N) * piximg;
                d.wp = q.d_wp; d.bias = q.d_bp; d.cout = q.cout; d.cout_pad = q.cout_pad;
                d.relu = 0; d.slope = (py == 1 && leaky) ? 0.1f : 0.f;
                d.raw = py == 0 ? tmp : nullptr;
                d.res = py == 1 ? tmp : nullptr;
                Op op;
                op.umma = true;
                op.lane = lane;
                std::string why;
                if (!fisr::build_split_conv(c->encode, d, c->num_sms, c->d_err, &op.conv, &why)) {
                    rc = fail(c, FISR_E_CUDA, "conv %s (stride 2, row phase %d): %s", name.c_str(), py, why.c_str());
                    return;
                }
                pl->ops.push_back(std::move(op));
                pl->umma_ops++;
            }
            return;
        }
        PwcConv a{};
        a.in = in.b; a.in_coff = in.coff; a.cin = p.cin_pad;
        a.out = outv.b; a.out_coff = outv.coff; a.cout = p.cout; a.out_f32 = out_f32;
        a.w = p.d_w; a.b = p.d_b;
        a.N = N; a.Hin = Hin; a.Win = Win; a.Hout = Hout; a.Wout = Wout;
        a.stride = stride; a.dil = dil;
        // tf 'same': total = (out - 1) s + 2 d + 1 - in, the smaller half first (0 before / 1 after for stride 2 on even sizes)
        a.pad_y = std::max((a.Hout - 1) * stride + 2 * dil + 1 - a.Hin, 0) / 2;
        a.pad_x = std::max((a.Wout - 1) * stride + 2 * dil + 1 - a.Win, 0) / 2;
        a.leaky = leaky ? 1 : 0; a.add = add; a.add_cs = 2;
        push([a](cudaStream_t st) { launch_conv3x3(a, st); });
    };
    auto conv = [&](const std::string& name, View in, View outv, int l_in, int l_out, int stride, int dil, bool leaky, float* out_f32 = nullptr,
                    const float* add = nullptr) {
        if (rc != FISR_OK) return;
        conv_p(P(name), name, in, outv, l_in, l_out, stride, dil, leaky, out_f32, add);
    };
    // ---- buffers
    Planes D[kLvls + 1] = {}, c2[kLvls + 1] = {}, tA[2][kLvls + 1] = {}, tB[2][kLvls + 1] = {};
    for (int l = kPredLvl; l <= kLvls; ++l) D[l] = planes(px(l), dense_cs(l));
    for (int l = 1; l <= kLvls; ++l) c2[l] = planes(px(l), pad8(kChann[l]));
    Planes c1c[kLvls + 1] = {};                                  // image 1's pyramid levels 2..5, compact (input of the next stride-2 conv);
    for (int l = 2; l < kLvls; ++l) c1c[l] = planes(px(l), kChann[l]);      // copied into their slot of D_l for the cascade
    for (int i = 0; i < 2; ++i) s2tmp[i] = f32(px(2) * kChann[2] + 512);
    Planes c1_1 = planes(px(1), pad8(kChann[1]));                // level 1 of image 1 feeds conv2a only
    Planes c1_6 = planes(px(kLvls), pad8(kChann[kLvls]));        // D_6 has no c1 slot (model_pwcnet.py:1549-1551)
    Planes F16 = planes(px(kPredLvl), 16);                       // output columns of the fused predict_flow / up_feat conv
    for (int img = 0; img < 2; ++img) {
        // scratch of one image's pyramid (the two pyramids run side by side): levels whose channel count is a multiple of 8 share two
        // buffers, the others get their own (their pad channels must stay zero: they are read by the TMA boxes of the next conv)
        Planes sA = planes(px(1), kChann[1]), sB = planes(px(1), kChann[1]);
        for (int l = 1; l <= kLvls; ++l) {
            if (kChann[l] % 8 == 0) { tA[img][l] = sA; tB[img][l] = sB; tA[img][l].cs = tB[img][l].cs = kChann[l]; }
            else { tA[img][l] = planes(px(l), pad8(kChann[l])); tB[img][l] = planes(px(l), pad8(kChann[l])); }
        }
    }
    Planes T0 = planes(px(kPredLvl), 128), T1 = planes(px(kPredLvl), 128);
    Planes warpbuf = planes(px(kPredLvl), 128);
    float* flow_raw = f32(px(kPredLvl) * 2);
    for (int l = kPredLvl; l <= kLvls; ++l) plan->flow[l] = f32(px(l) * 2);
    if (rc != FISR_OK) return rc;
    auto c1_view = [&](int l) -> View {
        if (l == 1) return View{c1_1, 0, kChann[1]};
        if (l == kLvls) return View{c1_6, 0, kChann[l]};
        return View{D[l], kActs + kCorrPad, kChann[l]};
    };
    // ---- feature pyramids (model_pwcnet.py:1012-1101): the Siamese extractor on both images
    for (int img = 0; img < 2; ++img) {
        View x{};
        lane = c->use_umma >= 3 ? img : 0;
        for (int l = 1; l <= kLvls; ++l) {
            const std::string p = "pwcnet/featpyr/conv" + std::to_string(l);
            const int f = kChann[l];
            const bool via_compact = img == 0 && l >= 2 && l < kLvls;
            View dst = via_compact ? View{c1c[l], 0, f} : (img == 0 ? c1_view(l) : View{c2[l], 0, f});
            View ta{tA[img][l], 0, f}, tb{tB[img][l], 0, f};
            if (l == 1) {        // straight from the fp32 image, whose pointer is bound per call
                const Param& pa = P(p + "a");
                const int which = img;
                const float *w1 = pa.d_w, *b1 = pa.d_b;
                const Planes o1 = tA[img][1];
                push([=](cudaStream_t st) { launch_first_conv(which ? pl->img2 : pl->img1, w1, b1, o1, N, H, W, st); });
            } else {
                conv(p + "a", x, ta, l - 1, l, 2, 1, true);
            }
            conv(p + "aa", ta, tb, l, l, 1, 1, true);
            conv(p + "b", tb, dst, l, l, 1, 1, true);
            if (via_compact) {
                const Planes srcp = c1c[l], dstp = D[l];
                const long long np = static_cast<long long>(px(l));
                push([=](cudaStream_t st) { launch_copy_channels(srcp, 0, dstp, kActs + kCorrPad, f, np, st); });
            }
            x = dst;
        }
    }
    lane = 0;
    // ---- coarse-to-fine cascade (model_pwcnet.py:1525-1593)
    for (int l = kLvls; l >= kPredLvl; --l) {
        const std::string sl = std::to_string(l);
        const int cs = dense_cs(l), h = H >> l, w = W >> l, C = kChann[l];
        const View c1v = c1_view(l);
        const Planes Dl = D[l], c2l = c2[l];
        bool fused_up = false;
        if (l == kLvls) {
            push([=](cudaStream_t st) { launch_cost_volume(c1v.b, c1v.coff, c2l, 0, C, Dl, kActs, N, h, w, st); });
        } else {
            const int uf = kActs + kCorrPad + C;           // up_flow slot
            const float scaler = 20.f / static_cast<float>(1 << l);
            Planes wb = warpbuf;
            wb.cs = pad8(C);
            push([=](cudaStream_t st) { launch_dense_warp(c2l, 0, C, Dl, uf, scaler, wb, N, h, w, st); });
            push([=](cudaStream_t st) { launch_cost_volume(c1v.b, c1v.coff, wb, 0, C, Dl, kActs, N, h, w, st); });
        }
        for (int k = 0; k < 5; ++k)
            conv("pwcnet/predict_flow/conv" + sl + "_" + std::to_string(k), View{Dl, dense_in_off(k), cs - dense_in_off(k)},
                 View{Dl, dense_off(k), kDense[k]}, l, l, 1, 1, true);
        if (c->use_umma && h >= 4 && w >= 4) {       // flow predictor + up_feat of the next level: one 16-column conv, then split
            if ((rc = build_fused(c, l)) != FISR_OK) return rc;
            Planes Fl = F16;
            conv_p(c->fused[l], "fused flow" + sl, View{Dl, 0, cs}, View{Fl, 0, 16}, l, l, 1, 1, false);
            const Planes Dn = l != kPredLvl ? D[l - 1] : Planes{};
            const int upf = l != kPredLvl ? kActs + kCorrPad + kChann[l - 1] + 2 : 0;
            push([=](cudaStream_t st) { launch_flow_upfeat_scatter(Fl, flow_raw, Dn, upf, N, h, w, st); });
            fused_up = true;
        } else {
            conv("pwcnet/predict_flow/flow" + sl, View{Dl, 0, cs}, View{}, l, l, 1, 1, false, flow_raw);
        }
        // context network (model_pwcnet.py:1453-1522): flow + dilated conv chain on upfeat
        View cur{Dl, 0, cs};
        for (int k = 0; k < 7; ++k) {
            const std::string nm = "pwcnet/ctxt/dc_conv" + sl + std::to_string(k + 1);
            if (k < 6) {
                Planes t = (k & 1) ? T1 : T0;
                t.cs = kCtxtF[k];
                View o{t, 0, kCtxtF[k]};
                conv(nm, cur, o, l, l, 1, kCtxtD[k], true);
                cur = o;
            } else {
                conv(nm, cur, View{}, l, l, 1, 1, false, plan->flow[l], flow_raw);
            }
        }
        if (l != kPredLvl) {
            const Param& pf = P("pwcnet/upsample/up_flow" + sl);
            const Param& pe = P("pwcnet/upsample/up_feat" + sl);
            const Planes Dn = D[l - 1];
            const int ufn = kActs + kCorrPad + kChann[l - 1];
            const float* fl = plan->flow[l];
            const float *wf = pf.d_w, *bf = pf.d_b, *we = pe.d_w, *be = pe.d_b;
            const int cine = pe.cin_pad;
            push([=](cudaStream_t st) { launch_deconv4x4s2(Planes{}, 0, fl, 2, 2, wf, bf, Dn, ufn, N, h, w, st); });
            if (!fused_up) push([=](cudaStream_t st) { launch_deconv4x4s2(Dl, 0, nullptr, 0, cine, we, be, Dn, ufn + 2, N, h, w, st); });
        }
    }
    {   // flow_pred = resize_bilinear(flow2, x4) * 4 (model_pwcnet.py:1588-1590)
        float* f2 = plan->flow[kPredLvl];
        const int h = H >> kPredLvl, w = W >> kPredLvl, S = 1 << kPredLvl;
        push([=](cudaStream_t st) { launch_resize_flow(f2, pl->out, N, h, w, S, static_cast<float>(S), st); });
    }
    if (rc != FISR_OK) return rc;
    PWC_TRY(c, cudaStreamSynchronize(c->stream));
    if (c->plans.size() >= 4) {          // a 1080p plan holds ~7 GB: keep at most four input sizes resident
        PWC_TRY(c, cudaDeviceSynchronize());
        c->plans.clear();
        c->last = nullptr;
    }
    *out = plan.get();
    c->plans[key] = std::move(plan);
    return FISR_OK;
}

}  // namespace

extern "C" {

int fisr_pwc_create(int device, fisr_pwc** out) {
    if (!out) return FISR_E_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(nullptr, FISR_E_CUDA, "no CUDA device (fisr_b200 has no CPU path)");
    if (device < 0 || device >= ndev) return fail(nullptr, FISR_E_INVALID, "device %d out of range", device);
    std::unique_ptr<fisr_pwc> c(new fisr_pwc());
    c->device = device;
    Guard guard(device);
    cudaFree(0);
    cudaDeviceProp prop;
    PWC_TRY(nullptr, cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(nullptr, FISR_E_CUDA, "fisr_b200 is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
    c->num_sms = prop.multiProcessorCount;
    {
        cudaDriverEntryPointQueryResult q;
        void* fn = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn)
            return fail(nullptr, FISR_E_CUDA, "cuTensorMapEncodeTiled unavailable");
        c->encode = reinterpret_cast<fisr::EncodeTiledFn>(fn);
    }
    if (const char* e = getenv("FISR_PWC_UMMA")) c->use_umma = atoi(e);
    PWC_TRY(nullptr, fisr::conv3x3_init());
    PWC_TRY(nullptr, init_kernels());
    PWC_TRY(nullptr, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (int i = 0; i < fisr_pwc::kSide; ++i) {
        PWC_TRY(nullptr, cudaStreamCreateWithFlags(&c->side[i], cudaStreamNonBlocking));
        PWC_TRY(nullptr, cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming));
    }
    PWC_TRY(nullptr, cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    PWC_TRY(nullptr, cudaMalloc(&c->d_err, sizeof(int)));
    PWC_TRY(nullptr, cudaMemset(c->d_err, 0, sizeof(int)));
    PWC_TRY(nullptr, cudaMallocHost(&c->h_err, sizeof(int)));
    *c->h_err = 0;
    const auto& inv = inventory();
    c->params.resize(inv.size());
    for (size_t i = 0; i < inv.size(); ++i) {
        Param& p = c->params[i];
        p.cin_pad = inv[i].cin + (inv[i].gap_at >= 0 ? kGap : 0);
        p.cout = inv[i].cout;
        const size_t wn = static_cast<size_t>(inv[i].transpose ? 16 : 9) * p.cin_pad * inv[i].cout;
        PWC_TRY(nullptr, cudaMalloc(&p.d_w, wn * sizeof(float)));
        PWC_TRY(nullptr, cudaMemset(p.d_w, 0, wn * sizeof(float)));
        PWC_TRY(nullptr, cudaMalloc(&p.d_b, inv[i].cout * sizeof(float)));
        PWC_TRY(nullptr, cudaMemset(p.d_b, 0, inv[i].cout * sizeof(float)));
        c->index[inv[i].name] = static_cast<int>(i);
    }
    PWC_TRY(nullptr, cudaDeviceSynchronize());
    *out = c.release();
    return FISR_OK;
}

void fisr_pwc_destroy(fisr_pwc* c) {
    if (!c) return;
    Guard guard(c->device);
    cudaDeviceSynchronize();
    c->plans.clear();
    for (auto& p : c->params) { cudaFree(p.d_w); cudaFree(p.d_b); cudaFree(p.d_wp); cudaFree(p.d_bp); }
    for (auto& p : c->fused) { cudaFree(p.d_w); cudaFree(p.d_b); cudaFree(p.d_wp); cudaFree(p.d_bp); }
    for (auto& q : c->s2) for (auto& p : q) { cudaFree(p.d_w); cudaFree(p.d_b); cudaFree(p.d_wp); cudaFree(p.d_bp); }
    for (auto& p : c->pk4) { cudaFree(p.d_w); cudaFree(p.d_b); cudaFree(p.d_wp); cudaFree(p.d_bp); }
    for (int i = 0; i < fisr_pwc::kSide; ++i) {
        if (c->side[i]) cudaStreamDestroy(c->side[i]);
        if (c->ev_join[i]) cudaEventDestroy(c->ev_join[i]);
    }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    cudaFree(c->d_err);
    if (c->h_err) cudaFreeHost(c->h_err);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

const char* fisr_pwc_last_error(const fisr_pwc* c) { return c ? c->err.c_str() : g_err.c_str(); }
int fisr_pwc_num_params(void) { return static_cast<int>(names().size()); }
const char* fisr_pwc_param_name(int i) { return (i >= 0 && i < (int)names().size()) ? names()[i].c_str() : nullptr; }
int fisr_pwc_param_shape(int i, int dims[4]) {
    const auto& inv = inventory();
    if (i < 0 || i >= 2 * (int)inv.size() || !dims) return FISR_E_INVALID;
    const PDef& d = inv[i / 2];
    if (i % 2) { dims[0] = d.cout; dims[1] = dims[2] = dims[3] = 1; return 1; }
    if (d.transpose) { dims[0] = 4; dims[1] = 4; dims[2] = d.cout; dims[3] = d.cin; }
    else { dims[0] = 3; dims[1] = 3; dims[2] = d.cin; dims[3] = d.cout; }
    return 4;
}

int fisr_pwc_set_param(fisr_pwc* c, const char* name, const float* h_data, size_t count) {
    if (!c || !name || !h_data) return FISR_E_INVALID;
    std::string s(name);
    const bool is_w = s.size() > 7 && s.substr(s.size() - 7) == "/kernel";
    const bool is_b = s.size() > 5 && s.substr(s.size() - 5) == "/bias";
    if (!is_w && !is_b) return fail(c, FISR_E_INVALID, "parameter name must end in /kernel or /bias: %s", name);
    auto it = c->index.find(s.substr(0, s.size() - (is_w ? 7 : 5)));
    if (it == c->index.end()) return fail(c, FISR_E_INVALID, "unknown parameter %s", name);
    const PDef& d = inventory()[it->second];
    Param& p = c->params[it->second];
    Guard guard(c->device);
    PWC_TRY(c, cudaDeviceSynchronize());
    if (is_b) {
        if (count != (size_t)d.cout) return fail(c, FISR_E_INVALID, "%s has %d elements, got %zu", name, d.cout, count);
        PWC_TRY(c, cudaMemcpy(p.d_b, h_data, count * sizeof(float), cudaMemcpyHostToDevice));
        p.h_b.assign(h_data, h_data + count);
        p.packed = false;
        for (auto& f : c->fused) f.packed = false;
        for (auto& f : c->s2) f[0].packed = f[1].packed = false;
        c->pk4[0].packed = c->pk4[1].packed = false;
        return FISR_OK;
    }
    const size_t taps = d.transpose ? 16 : 9, expect = taps * d.cin * d.cout;
    if (count != expect) return fail(c, FISR_E_INVALID, "%s has %zu elements, got %zu", name, expect, count);
    // kernel layout with the kGap zero rows of the padded cost-volume slot: reference input channel i sits at i (+kGap past the gap)
    std::vector<float> packed(taps * p.cin_pad * d.cout, 0.f);
    auto slot = [&](int ci) { return (d.gap_at >= 0 && ci >= d.gap_at) ? ci + kGap : ci; };
    for (size_t t = 0; t < taps; ++t)
        for (int ci = 0; ci < d.cin; ++ci)
            for (int co = 0; co < d.cout; ++co) {
                if (d.transpose) packed[(t * d.cout + co) * p.cin_pad + slot(ci)] = h_data[(t * d.cout + co) * d.cin + ci];      // [4,4,out,in]
                else packed[(t * p.cin_pad + slot(ci)) * d.cout + co] = h_data[(t * d.cin + ci) * d.cout + co];                  // [3,3,in,out]
            }
    PWC_TRY(c, cudaMemcpy(p.d_w, packed.data(), packed.size() * sizeof(float), cudaMemcpyHostToDevice));
    p.h_w.assign(h_data, h_data + count);
    p.packed = false;
    for (auto& f : c->fused) f.packed = false;
    for (auto& f : c->s2) f[0].packed = f[1].packed = false;
    c->pk4[0].packed = c->pk4[1].packed = false;
    return FISR_OK;
}

int fisr_pwc_forward(fisr_pwc* c, const float* d_img1, const float* d_img2, int N, int H, int W, float* d_flow, void* stream) {
    if (!c || !d_img1 || !d_img2 || !d_flow) return FISR_E_INVALID;
    Guard guard(c->device);
    Plan* plan = nullptr;
    if (*c->h_err) return fail(c, FISR_E_CUDA, "a conv kernel of an earlier forward timed out on a barrier (code %d)", *c->h_err);
    for (auto& p : c->params)            // parameters changed since the plan was built: re-pack the operand planes in place
        if (p.d_wp && !p.packed) { const int rp = ensure_packed(c, p); if (rp != FISR_OK) return rp; }
    for (int l = kPredLvl; l <= kLvls; ++l) {
        if (c->fused[l].d_wp && !c->fused[l].packed) { const int rp = build_fused(c, l); if (rp != FISR_OK) return rp; }
        if (c->s2[l][0].d_wp && !(c->s2[l][0].packed && c->s2[l][1].packed)) { const int rp = build_stride2(c, l); if (rp != FISR_OK) return rp; }
    }
    for (int k = 0; k < 2; ++k)
        if (c->pk4[k].d_wp && !c->pk4[k].packed) { const int rp = build_packed4(c, k, k ? "pwcnet/featpyr/conv1b" : "pwcnet/featpyr/conv1aa"); if (rp != FISR_OK) return rp; }
    int rc = build_plan(c, N, H, W, &plan);
    if (rc != FISR_OK) return rc;
    cudaStream_t st0 = stream ? static_cast<cudaStream_t>(stream) : c->stream;
    if (st0 != c->stream) PWC_TRY(c, cudaStreamSynchronize(c->stream));      // operand packing / buffer clears ran on the context stream
    plan->img1 = d_img1; plan->img2 = d_img2; plan->out = d_flow;
    const bool fan_out = c->use_umma >= 3;
    bool forked = false;
    for (size_t i = 0; i < plan->ops.size(); ++i) {
        Op& op = plan->ops[i];
        if (op.lane == 1 && !forked) {          // first op of image 2's pyramid: it depends on nothing before it in this forward
            PWC_TRY(c, cudaEventRecord(c->ev_fork, st0));
            PWC_TRY(c, cudaStreamWaitEvent(c->side[0], c->ev_fork, 0));
            forked = true;
        }
        if (op.lane == 0 && forked) {           // first op of the cascade: needs both pyramids
            PWC_TRY(c, cudaEventRecord(c->ev_join[0], c->side[0]));
            PWC_TRY(c, cudaStreamWaitEvent(st0, c->ev_join[0], 0));
            forked = false;
        }
        cudaStream_t st = op.lane == 1 ? c->side[0] : st0;
        if (!op.umma) { op.fn(st); continue; }
        if (op.group == 0 || !fan_out) { PWC_TRY(c, fisr::launch_conv3x3(op.conv, c->num_sms, st)); continue; }
        // a group of independent launches: fork to the side streams, round robin, join
        size_t end = i;
        while (end < plan->ops.size() && plan->ops[end].umma && plan->ops[end].group == op.group) ++end;
        PWC_TRY(c, cudaEventRecord(c->ev_fork, st));
        for (int k = 0; k < fisr_pwc::kSide; ++k) PWC_TRY(c, cudaStreamWaitEvent(c->side[k], c->ev_fork, 0));
        for (size_t j = i; j < end; ++j) {
            const int lane = static_cast<int>((j - i) % (fisr_pwc::kSide + 1));
            PWC_TRY(c, fisr::launch_conv3x3(plan->ops[j].conv, c->num_sms, lane == 0 ? st : c->side[lane - 1]));
        }
        for (int k = 0; k < fisr_pwc::kSide; ++k) {
            PWC_TRY(c, cudaEventRecord(c->ev_join[k], c->side[k]));
            PWC_TRY(c, cudaStreamWaitEvent(st, c->ev_join[k], 0));
        }
        i = end - 1;
    }
    PWC_TRY(c, cudaMemcpyAsync(c->h_err, c->d_err, sizeof(int), cudaMemcpyDeviceToHost, st0));
    c->launches += static_cast<long long>(plan->ops.size());
    c->last = plan;
    PWC_TRY(c, cudaGetLastError());
    return FISR_OK;
}

int fisr_pwc_prepare_pair(fisr_pwc* c, const void* d_frame0, const void* d_frame1, int src_kind, const double* h_yuv2rgb, int h, int w, int scale,
                          float* d_img1, float* d_img2, void* stream) {
    if (!c || !d_frame0 || !d_frame1 || !d_img1 || !d_img2) return FISR_E_INVALID;
    if (h < 2 || w < 2 || scale < 1 || (src_kind != 0 && src_kind != 1) || (src_kind == 0 && !h_yuv2rgb))
        return fail(c, FISR_E_INVALID, "fisr_pwc_prepare_pair: %d x %d, scale %d, source kind %d", h, w, scale, src_kind);
    Guard guard(c->device);
    PrepConst k{};
    if (src_kind == 0) { memcpy(k.T, h_yuv2rgb, 9 * sizeof(double)); memcpy(k.off, h_yuv2rgb + 9, 3 * sizeof(double)); }
    const int oh = h * scale, ow = w * scale;
    k.ry = static_cast<double>(h) / static_cast<double>(oh);
    k.rx = static_cast<double>(w) / static_cast<double>(ow);
    const int Hp = (oh + 63) / 64 * 64, Wp = (ow + 63) / 64 * 64;
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : c->stream;
    launch_prepare_pair(d_frame0, d_frame1, src_kind == 0, d_img1, d_img2, h, w, oh, ow, Hp, Wp, k, st);
    c->launches++;
    PWC_TRY(c, cudaGetLastError());
    return FISR_OK;
}

int fisr_pwc_finish_flow(fisr_pwc* c, const float* d_flow, int N, int Hp, int Wp, int h0, int w0, int h_out, int w_out, const double* h_wy, int ry,
                         const double* h_wx, int rx, double scale, float* d_out, void* stream) {
    if (!c || !d_flow || !d_out) return FISR_E_INVALID;
    if (N < 1 || h0 < 1 || w0 < 1 || h0 > Hp || w0 > Wp || h_out < 1 || w_out < 1 || ry < 0 || rx < 0 || ry > kMaxGaussRadius || rx > kMaxGaussRadius ||
        (ry > 0 && !h_wy) || (rx > 0 && !h_wx) || ry >= h0 || rx >= w0 || !(scale > 0))
        return fail(c, FISR_E_INVALID, "fisr_pwc_finish_flow: crop %d x %d of %d x %d -> %d x %d, radii %d / %d", h0, w0, Hp, Wp, h_out, w_out, ry, rx);
    Guard guard(c->device);
    FinishConst k{};
    k.wy[0] = k.wx[0] = 1.0;
    k.ry = ry; k.rx = rx;
    if (ry > 0) memcpy(k.wy, h_wy, (ry + 1) * sizeof(double));
    if (rx > 0) memcpy(k.wx, h_wx, (rx + 1) * sizeof(double));
    k.ratio_y = static_cast<double>(h0) / static_cast<double>(h_out);
    k.ratio_x = static_cast<double>(w0) / static_cast<double>(w_out);
    k.scale = scale;
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : c->stream;
    launch_finish_flow(d_flow, N, Hp, Wp, h0, w0, h_out, w_out, d_out, k, st);
    c->launches++;
    PWC_TRY(c, cudaGetLastError());
    return FISR_OK;
}

int fisr_pwc_debug_flow(fisr_pwc* c, int lvl, float* h_dst, size_t count) {
    if (!c || !h_dst || lvl < kPredLvl || lvl > kLvls) return FISR_E_INVALID;
    if (!c->last) return fail(c, FISR_E_INVALID, "no forward has run yet");
    Guard guard(c->device);
    const size_t n = static_cast<size_t>(c->last->N) * (c->last->H >> lvl) * (c->last->W >> lvl) * 2;
    if (count != n) return fail(c, FISR_E_INVALID, "flow%d has %zu elements, got %zu", lvl, n, count);
    PWC_TRY(c, cudaDeviceSynchronize());
    if (*c->h_err) return fail(c, FISR_E_CUDA, "a conv kernel timed out on a barrier (code %d)", *c->h_err);
    PWC_TRY(c, cudaMemcpy(h_dst, c->last->flow[lvl], n * sizeof(float), cudaMemcpyDeviceToHost));
    return FISR_OK;
}

long long fisr_pwc_launch_count(const fisr_pwc* c) { return c ? c->launches : 0; }

}  // extern "C"
