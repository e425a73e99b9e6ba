// C ABI (include/fisr_b200.h): context, parameter store, forward plan (buffers + TMA descriptors + launch list,
// replayed as a CUDA graph) for FISRnet.model (FISRnet.py:73-173), the tiled window driver
// (FISRnet.py:994-1065) and the flow warp.  No torch types; PyTorch only hands in raw device pointers.
#include <algorithm>
#include <cmath>
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
#include "common.cuh"
#include "conv_umma.h"
#include "train_kernels.h"
#include "wgrad_umma.h"

using namespace fisr;

namespace {


constexpr int CH = 64;       // FISRnet.py:74
constexpr int IN_CH = 29;    // FISRnet.py:287

struct ParamDef {
    std::string name;        // conv name without /w, /b
    int cin, cout;
};

// Creation order of the variables in FISRnet.model (FISRnet.py:78-171; block scopes ops.py:40,49,60,68).
std::vector<ParamDef> build_inventory() {
    std::vector<ParamDef> v;
    auto res_block = [&](const std::string& p, int c) {
        v.push_back({p + "/conv/0", c, c});
        v.push_back({p + "/conv/1", c, c});
    };
    auto enc = [&](const std::string& p, int c1, int c) {
        v.push_back({p + "/conv/0", c1, c});
        res_block(p + "/res_block/0", c);
        res_block(p + "/res_block/1", c);
    };
    auto dec = [&](const std::string& p, int c1, int c) {
        v.push_back({p + "/resize", c1, c});
        v.push_back({p + "/conv/0", 2 * c, c});
        res_block(p + "/res_block/0", c);
        res_block(p + "/res_block/1", c);
    };
    auto head = [&](const std::string& p, int cout) {
        v.push_back({p + "/conv/0", CH, CH});
        res_block(p + "/res_block/0", CH);
        v.push_back({p + "/conv/1", CH, CH * 4});
        v.push_back({p + "/conv/2", CH, cout});
    };
    for (int lvl = 1; lvl <= 3; ++lvl) {
        const std::string p = "FISRnet/level_" + std::to_string(lvl);
        enc(p + "/enc/level_0", lvl == 1 ? IN_CH : IN_CH + 9, CH);
        enc(p + "/enc/level_1", CH, CH * 2);
        enc(p + "/enc/level_2", CH * 2, CH * 4);
        v.push_back({p + "/bottleneck/conv/0", CH * 4, CH * 8});
        res_block(p + "/bottleneck/res_block/0", CH * 8);
        dec(p + "/dec/level_2", CH * 8, CH * 4);
        dec(p + "/dec/level_1", CH * 4, CH * 2);
        dec(p + "/dec/level_0", CH * 2, CH);
        head(p + "/FI-SR", 6);
        head(p + "/SR", 3);
    }
    return v;
}

const std::vector<ParamDef>& inventory() {
    static const std::vector<ParamDef> inv = build_inventory();
    return inv;
}
const std::vector<std::string>& param_names() {
    static std::vector<std::string> names;
    if (names.empty())
        for (const auto& d : inventory()) {
            names.push_back(d.name + "/w");
            names.push_back(d.name + "/b");
        }
    return names;
}

struct ConvParam {
    int cin = 0, cout = 0, KB = 0, cout_pad = 0;
    float* d_w = nullptr;        // fp32 HWIO master copy
    float* d_b = nullptr;        // fp32 [cout_pad], zero padded
    __half* d_wp = nullptr;      // packed (hi, lo) operand planes
    bool packed = false;
    float *m_w = nullptr, *v_w = nullptr, *m_b = nullptr, *v_b = nullptr;   // Adam slots (allocated on first use)
    // training: operand planes of the transposed / rotated filter (dgrad) and the gradients of the last backward
    int cin_pad = 0, OBk = 0;    // Cin rounded up to 64 (dgrad N extent), 64-channel blocks of Cout (dgrad K blocks)
    __half* d_wpT = nullptr;
    bool packedT = false;
    float *g_w = nullptr, *g_b = nullptr;
    // conv/2 heads under f16f8: the conv evaluated at input resolution with depth_to_space folded into the weights
    // (aux_kernels.cu, expand_ps_weights_kernel): fp32 [3,3,256,4*cout] + bias [32], packed planes, cout_pad 16 or 32
    bool head2 = false;
    float *d_wps = nullptr, *d_bps = nullptr;
    __half* d_wpps = nullptr;
    int ps_cout_pad = 0;
};

// Buffers of one forward level that the backward pass reads again (activations = ReLU masks and wgrad operands).
struct ResRec { ActBuf a1, a2, a3; };            // two_res_blocks: relu(conv0(a0)), relu(n1), relu(conv0(a2))
struct EncRec { std::string p; ActBuf x; int x_cs = 0; int c = 0, H = 0, W = 0; ActBuf a0, cat, pooled; ResRec rb; };
struct BottRec { std::string p; ActBuf x, a0, a1, out; int H = 0, W = 0; };
struct DecRec { std::string p; ActBuf x_in, up, cat, a0, out; ResRec rb; int c1 = 0, c = 0, H = 0, W = 0; };
struct HeadRec { std::string p; ActBuf x, a0, a1, a2, shuf; int cout = 0; };
struct LevelRec { int lvl = 0, N = 0, H = 0, W = 0; ActBuf in; EncRec enc[3]; BottRec bott; DecRec dec[3]; HeadRec head[2]; };

struct BwdOp {
    std::function<int(cudaStream_t)> run;
    std::string name;
    double flops = 0;
    int launches = 1;
};

struct DebugTensor {
    const float* raw = nullptr;
    int N = 0, H = 0, W = 0, C = 0;
};

enum OpKind { OP_CONV, OP_UPSAMPLE, OP_PRED2NEXT };
struct Op {
    OpKind kind;
    ConvLaunch conv;
    ActBuf in, out;
    const float* pred = nullptr;   // OP_PRED2NEXT source
    int N, H, W, C, cs, coff;
    double flops;            // algorithmic conv FLOPs (0 for pool / upsample)
    double bytes;            // algorithmic HBM bytes of the launch
    char name[96];
};

struct Plan {
    int N = 0, H = 0, W = 0, planes = 0;
    std::vector<void*> allocs;
    size_t bytes = 0;
    std::vector<Op> ops;
    ActBuf in_lvl[3];
    float* pred[3] = {nullptr, nullptr, nullptr};
    int pred_cs = 9;                 // floats per prediction pixel: 9, or 12 (padded groups) under f16f8
    float* pred9[3] = {nullptr, nullptr, nullptr};      // compact copies for fisr_forward when pred_cs == 12
    std::map<std::string, DebugTensor> debug;
    cudaGraphExec_t graph = nullptr;
    double flops = 0, eff_weighted = 0;
    // training (built on demand by ensure_backward)
    LevelRec rec[3];
    std::vector<BwdOp> bwd;
    bool has_bwd = false;
    int B = 0;                       // training batch (N = 4B passes)
    float loss_scale = 1.f;
    ActBuf gp[3];                    // dy operand of the conv/2 heads per level
    LossLambdas lam{1.f, 1.f, 0.1f, 1.f, 0.1f, 1.f};
    const float* label = nullptr;    // set per call
    float* wg_partial = nullptr;     // split-K slots of the wgrad kernel (weights + bias, sized for the largest layer)
    ~Plan() {
        if (graph) cudaGraphExecDestroy(graph);
        for (void* p : allocs) cudaFree(p);
    }
};

}  // namespace

struct fisr_ctx {
    int device = 0;
    int num_sms = 148;
    int planes = 2;
    bool use_graph = true;
    cudaStream_t stream = nullptr;
    EncodeTiledFn encode = nullptr;
    std::vector<ConvParam> params;
    std::map<std::string, int> conv_index;
    std::map<std::string, std::unique_ptr<Plan>> plans;
    int* d_err = nullptr;
    float* d_lut255 = nullptr;
    long long launches = 0;
    std::string err;
    // staging for the host entry points
    void* stage[8] = {nullptr};
    size_t stage_bytes[8] = {0};
    Plan* last_plan = nullptr;
    // pipelined host path (fisr_window_submit / fisr_window_wait): copies run on their own streams
    struct HostSlot {
        void* in[3] = {nullptr, nullptr, nullptr};
        size_t in_bytes[3] = {0, 0, 0};
        void* out = nullptr;
        size_t out_bytes = 0;
        cudaEvent_t h2d_done = nullptr, compute_done = nullptr, d2h_done = nullptr;
        int* h_err = nullptr;        // pinned copy of the kernel error flag, filled behind the canvas copy
        bool busy = false, used = false;
    } slots[2];
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    long long adam_t = 0;        // Adam step counter (global_step of FISRnet.py:232,491)
    float* d_scalars = nullptr;  // 11 loss scalars
    float* d_zero_bias = nullptr;   // 512 zeros: the dgrad launches of the conv kernel add no bias
    unsigned* d_gmax = nullptr;     // max |gradient| bits of the last backward (overflow check of the loss scale)
    float loss_scale_override = 0.f;
    bool wgrad_exact = false;       // multiply by x's lo plane in every wgrad launch (fisr_set_wgrad_exact)
    std::vector<MtTensor> mt_host;  // multi-tensor optimiser tables (train_kernels.h)
    MtTensor* d_mt = nullptr;
    MtPack* d_mt_pack = nullptr;
    size_t mt_pack_cap = 0;
};

namespace {

std::string g_create_error;

int fail(fisr_ctx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_create_error = buf;
    return code;
}

#define CUDA_TRY(ctx, expr)                                                                               \
    do {                                                                                                  \
        cudaError_t e__ = (expr);                                                                         \
        if (e__ != cudaSuccess)                                                                           \
            return fail(ctx, FISR_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

struct Guard {   // makes ctx->device current for the duration of a call
    int prev = -1;
    explicit Guard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~Guard() { if (prev >= 0) cudaSetDevice(prev); }
};

int ensure_stage(fisr_ctx* ctx, int slot, size_t bytes) {
    if (ctx->stage_bytes[slot] >= bytes) return FISR_OK;
    if (ctx->stage[slot]) cudaFree(ctx->stage[slot]);
    ctx->stage[slot] = nullptr;
    ctx->stage_bytes[slot] = 0;
    CUDA_TRY(ctx, cudaMalloc(&ctx->stage[slot], bytes));
    ctx->stage_bytes[slot] = bytes;
    return FISR_OK;
}

int check_kernel_error(fisr_ctx* ctx) {
    int h = 0;
    CUDA_TRY(ctx, cudaMemcpyAsync(&h, ctx->d_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (h != 0) {
        cudaMemset(ctx->d_err, 0, sizeof(int));
        return fail(ctx, FISR_E_KERNEL, "conv kernel pipeline time-out, barrier code %d", h);
    }
    return FISR_OK;
}

int ensure_packed(fisr_ctx* ctx, ConvParam& p, cudaStream_t st) {
    if (p.packed) return FISR_OK;
    // f16f8: the first conv of levels 2 and 3 (29 + 9 input channels) reads its 9 prediction channels from slot kPredSlot
    const bool pred_in = ctx->planes == 3 && p.cin == IN_CH + 9;
    launch_prep_weights(p.d_w, p.d_wp, p.cin, p.cout, p.KB, p.cout_pad, ctx->planes, st, pred_in ? IN_CH : 1 << 30,
                        pred_in ? kPredSlot - IN_CH : 0);
    ctx->launches++;
    if (p.head2 && ctx->planes == 3) {
        if (!p.d_wps) {
            CUDA_TRY(ctx, cudaMalloc(&p.d_wps, static_cast<size_t>(9) * 256 * 4 * p.cout * sizeof(float)));
            CUDA_TRY(ctx, cudaMalloc(&p.d_bps, 32 * sizeof(float)));
            CUDA_TRY(ctx, cudaMemset(p.d_bps, 0, 32 * sizeof(float)));
            CUDA_TRY(ctx, cudaMalloc(&p.d_wpps, static_cast<size_t>(2) * 4 * 9 * p.ps_cout_pad * 64 * sizeof(__half)));
        }
        launch_expand_ps_weights(p.d_w, p.d_b, p.d_wps, p.d_bps, p.cout, st);
        launch_prep_weights(p.d_wps, p.d_wpps, 256, 4 * p.cout, 4, p.ps_cout_pad, 3, st);
        ctx->launches += 2;
    }
    CUDA_TRY(ctx, cudaGetLastError());
    p.packed = true;
    return FISR_OK;
}

// ---------------------------------------------------------------- plan builder
struct Builder {
    fisr_ctx* ctx;
    Plan* plan;
    int rc = FISR_OK;

    void* alloc(size_t bytes, bool zero) {
        void* p = nullptr;
        if (rc != FISR_OK) return nullptr;
        bytes = (bytes + 1023) / 1024 * 1024;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) {
            rc = fail(ctx, FISR_E_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
            return nullptr;
        }
        if (zero) cudaMemsetAsync(p, 0, bytes, ctx->stream);
        plan->allocs.push_back(p);
        plan->bytes += bytes;
        return p;
    }
    ActBuf act(int N, int H, int W, int C, bool zero = false) {
        ActBuf b;
        const size_t elems = static_cast<size_t>(N) * H * W * C;
        b.plane = (elems + 511) / 512 * 512;
        b.p = static_cast<__half*>(alloc(b.plane * act_planes(plan->planes) * sizeof(__half), zero));
        return b;
    }
    float* f32(int N, int H, int W, int C) { return static_cast<float*>(alloc(static_cast<size_t>(N) * H * W * C * 4, false)); }

    bool encode_act(CUtensorMap* tm, const __half* base, int cs, int N, int H, int W, int P, int rows) {
        cuuint64_t dims[4] = {(cuuint64_t)cs, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)cs * 2, (cuuint64_t)W * cs * 2, (cuuint64_t)H * W * cs * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)P, (cuuint32_t)rows, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = ctx->encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half*>(base), dims, strides, box, es,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            rc = fail(ctx, FISR_E_CUDA, "cuTensorMapEncodeTiled(act %dx%dx%dx%d box %dx%d) failed: %d", N, H, W, cs, P, rows, (int)r);
            return false;
        }
        return true;
    }
    // TMA-store maps of an activation buffer (f16f8): fp16 plane as [N][H][W][cs] with boxes of 16 channels x 8 px x 4 rows, and
    // the 8-bit rows as [N][H][W][2 * cs / 64 half-rows][64 bytes] with boxes of 16 bytes x (lo, hi half-row) x 8 px x 4 rows
    bool encode_out(CUtensorMap* tm16, CUtensorMap* tm8, ActBuf buf, int cs, int N, int H, int W) {
        {
            cuuint64_t dims[4] = {(cuuint64_t)cs, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
            cuuint64_t strides[3] = {(cuuint64_t)cs * 2, (cuuint64_t)W * cs * 2, (cuuint64_t)H * W * cs * 2};
            cuuint32_t box[4] = {16, 8, 4, 1};
            cuuint32_t es[4] = {1, 1, 1, 1};
            CUresult r = ctx->encode(tm16, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, buf.p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { rc = fail(ctx, FISR_E_CUDA, "cuTensorMapEncodeTiled(out fp16 %dx%dx%dx%d) failed: %d", N, H, W, cs, (int)r); return false; }
        }
        {
            cuuint64_t dims[5] = {64, (cuuint64_t)(2 * (cs / 64)), (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
            cuuint64_t strides[4] = {64, (cuuint64_t)cs * 2, (cuuint64_t)W * cs * 2, (cuuint64_t)H * W * cs * 2};
            cuuint32_t box[5] = {16, 2, 8, 4, 1};
            cuuint32_t es[5] = {1, 1, 1, 1, 1};
            CUresult r = ctx->encode(tm8, CU_TENSOR_MAP_DATA_TYPE_UINT8, 5, buf.p + buf.plane, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { rc = fail(ctx, FISR_E_CUDA, "cuTensorMapEncodeTiled(out 8-bit %dx%dx%dx%d) failed: %d", N, H, W, cs, (int)r); return false; }
        }
        return true;
    }
    bool encode_w(CUtensorMap* tm, const __half* base, size_t rows, int NT) {
        cuuint64_t dims[2] = {64, (cuuint64_t)rows};
        cuuint64_t strides[1] = {128};
        cuuint32_t box[2] = {64, (cuuint32_t)NT};
        cuuint32_t es[2] = {1, 1};
        CUresult r = ctx->encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), dims, strides, box, es,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            rc = fail(ctx, FISR_E_CUDA, "cuTensorMapEncodeTiled(weights rows %zu box %d) failed: %d", rows, NT, (int)r);
            return false;
        }
        return true;
    }

    struct ConvOut {
        const float* res = nullptr;
        int res_cs = 0;
        float* raw = nullptr;
        int raw_cs = 0, raw_off0 = 0, raw_off1 = 0, raw_split = 0;
        ActBuf act;
        int act_cs = 0, act_off0 = 0, act_off1 = 0, act_split = 0;
        bool relu = true, d2s = false, scalar = false;
        bool ps = false;          // conv/2 head at input resolution: output columns are (sub-pixel, channel) pairs
        ActBuf pool;              // fused 2x2 max-pool of the activation output (closing conv of an encoder level), pool_cs channels
        int pool_cs = 0;
        ActBuf mask;              // dgrad: ReLU gate (hi plane of a forward activation), mask_cs channels per pixel
        int mask_cs = 0, mask_off = 0;
        bool s2d = false;         // dgrad of conv/2: store space-to-depth into a 256-channel buffer
    };
    // Operand view of a parameter: forward (w) or data-gradient (rotated transpose) planes.
    struct WView {
        const __half* wp;
        int KB, cout_pad, cout, cin;
        const float* bias;
    };
    WView fwd_view(const ConvParam& p) const { return WView{p.d_wp, p.KB, p.cout_pad, p.cout, p.cin, p.d_b}; }
    WView ps_view(const ConvParam& p) const { return WView{p.d_wpps, 4, p.ps_cout_pad, 4 * p.cout, 256, p.d_bps}; }
    WView bwd_view(const ConvParam& p) const { return WView{p.d_wpT, p.OBk, p.cin_pad, p.cin, p.cout, ctx->d_zero_bias}; }

    // One conv launch: input = channels [cin_off, cin_off + KB*64) of `in` (cs channels, N x H x W).
    void conv(const ConvParam& p, ActBuf in, int in_cs, int cin_off, int N, int H, int W, const ConvOut& o,
              const std::string& name) {
        Op op{};
        if (!make_conv(fwd_view(p), in, in_cs, cin_off, N, H, W, o, name, &op)) return;
        if (o.raw && !o.scalar) plan->debug[name] = DebugTensor{o.raw, N, H, W, o.raw_cs};
        plan->flops += op.flops;
        plan->eff_weighted += op.flops * op.conv.efficiency;
        plan->ops.push_back(op);
    }
    // conv/2 head with folded depth_to_space (planes == 3): 256 -> 4 * cout at input resolution, writes [N,2H,2W,*] outputs
    void conv_ps(const ConvParam& p, ActBuf in, int N, int H, int W, const ConvOut& o, const std::string& name) {
        Op op{};
        if (!make_conv(ps_view(p), in, 256, 0, N, H, W, o, name, &op)) return;
        ConvArgs& a = op.conv.args;
        a.ps_cout = p.cout;
        for (int g = 0; g < 4; ++g) {          // source group g = 2i'+j' is read by R-taps dy in {0,1} (i' = 0) or {-1,0} (i' = 1), same in x
            unsigned m = 0;
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    const bool ay = (g >> 1) == 0 ? (dy >= 0) : (dy <= 0), ax = (g & 1) == 0 ? (dx >= 0) : (dx <= 0);
                    if (ay && ax) m |= 1u << ((dy + 1) * 3 + dx + 1);
                }
            a.tapmask[g] = m;
        }
        op.flops = 2.0 * 9 * 64 * p.cout * static_cast<double>(2 * H) * (2 * W) * N;      // algorithmic: the 64 -> cout conv at 2R
        plan->flops += op.flops;
        plan->eff_weighted += op.flops * op.conv.efficiency;
        plan->ops.push_back(op);
    }
    bool make_conv(const WView& p, ActBuf in, int in_cs, int cin_off, int N, int H, int W, const ConvOut& o,
                   const std::string& name, Op* out) {
        if (rc != FISR_OK) return false;
        Op& op = *out;
        op.kind = OP_CONV;
        ConvLaunch& L = op.conv;
        if (!plan_conv_geometry(H, W, N, p.cout_pad, plan->planes, ctx->num_sms, &L, p.KB)) {
            rc = fail(ctx, FISR_E_INVALID, "no tile geometry for conv %s (%dx%d)", name.c_str(), H, W);
            return false;
        }
        ConvArgs& a = L.args;
        a.bias = p.bias;
        a.res = o.res; a.res_cs = o.res_cs;
        a.out_raw = o.raw; a.raw_cs = o.raw_cs; a.raw_off0 = o.raw_off0; a.raw_off1 = o.raw_off1; a.raw_split = o.raw_split;
        a.out_act = o.act.p; a.act_plane = o.act.plane;
        a.act_cs = o.act_cs; a.act_off0 = o.act_off0; a.act_off1 = o.act_off1; a.act_split = o.act_split;
        a.act_relu = o.relu; a.act_d2s = o.d2s; a.scalar_out = o.scalar;
        a.mask = o.mask.p; a.mask_cs = o.mask_cs; a.mask_off = o.mask_off;
        a.pool_out = o.pool.p; a.pool_plane = o.pool.plane; a.pool_cs = o.pool_cs;
        a.err = ctx->d_err;
        a.N = N; a.H = H; a.W = W;
        L.epi = o.scalar ? 0 : ((o.res ? 1 : 0) | (o.raw ? 2 : 0) | (o.d2s ? 4 : 0) | (o.mask.p ? 8 : 0) | (o.s2d ? 16 : 0));
        if (!o.scalar && !o.act.p) { rc = fail(ctx, FISR_E_INVALID, "conv %s: the wide epilogue always writes an activation", name.c_str()); return false; }
        if (o.pool.p && (L.epi != 1 || (H & 1) || (W & 1))) { rc = fail(ctx, FISR_E_INVALID, "conv %s: the fused max-pool needs the residual epilogue and even H, W", name.c_str()); return false; }
        if (o.s2d && (p.cout_pad != 64 || (H & 1) || (W & 1))) { rc = fail(ctx, FISR_E_INVALID, "conv %s: space-to-depth needs 64 output channels and even H, W", name.c_str()); return false; }
        {   // the epilogue indexes with 32-bit element offsets
            const double px = static_cast<double>(N) * H * W * ((o.d2s || o.ps) ? 4 : 1);
            const int cs_max = std::max(std::max(std::max(o.raw_cs, o.res_cs), o.act_cs), o.mask_cs);
            if (px * cs_max >= 4294967296.0) { rc = fail(ctx, FISR_E_INVALID, "conv %s: %d x %d x %d x %d exceeds 32-bit offsets; split the batch", name.c_str(), N, H, W, cs_max); return false; }
        }
        a.cin_off = cin_off; a.KB = p.KB; a.cout = p.cout;
        {   // padded input channels beyond the last real one are zeros: skip their MMAs (f16f8 level-2/3 inputs: 9 channels at slot 32)
            const int cin_slots = (plan->planes == 3 && p.cin == IN_CH + 9 && p.KB == 1) ? kPredSlot + 9 : p.cin;
            a.ksteps_last = (cin_slots - (p.KB - 1) * 64 + 15) / 16;
        }
        if (a.ksteps_last < 1 || a.ksteps_last > 4) a.ksteps_last = 4;
        if (!encode_act(&L.tmA_hi, in.p, in_cs, N, H, W, a.P, a.TH + 2)) return false;
        if (!encode_act(&L.tmA_lo, in.p + (plan->planes >= 2 ? in.plane : 0), in_cs, N, H, W, a.P, a.TH + 2)) return false;
        L.tma_out = false;
        a.tma_out = 0;
        memset(&L.tmO_hi, 0, sizeof L.tmO_hi);
        memset(&L.tmO_8, 0, sizeof L.tmO_8);
        if (plan->planes == 3 && L.epi == 0 && !o.scalar && o.act.p && o.act_cs % 64 == 0) {
            if (!encode_out(&L.tmO_hi, &L.tmO_8, o.act, o.act_cs, N, H, W)) return false;
            L.tma_out = true;
            a.tma_out = 1;
        }
        if (!encode_w(&L.tmB, p.wp, static_cast<size_t>(act_planes(plan->planes)) * p.KB * 9 * p.cout_pad, L.pair ? L.NT / 2 : L.NT)) return false;
        op.flops = 2.0 * 9 * p.cin * p.cout * static_cast<double>(H) * W * N;
        {   // algorithmic HBM bytes: input + weights once, every output / residual once
            const double px = static_cast<double>(N) * H * W, eb = 2.0 * act_planes(plan->planes);
            op.bytes = px * p.KB * 64 * eb + 9.0 * p.KB * 64 * p.cout_pad * eb + (o.res ? px * p.cout * 4 : 0) +
                       (o.raw ? px * p.cout * 4 : 0) + (o.act.p ? px * p.cout * eb : 0) + (o.mask.p ? px * p.cout * 2 : 0) +
                       (o.pool.p ? px / 4 * p.cout * eb : 0);
        }
        snprintf(op.name, sizeof op.name, "%s", name.c_str());
        return true;
    }
    void upsample(ActBuf in, ActBuf out, int N, int h, int w, int C) {
        Op op{};
        op.kind = OP_UPSAMPLE; op.in = in; op.out = out; op.N = N; op.H = h; op.W = w; op.C = C;
        op.bytes = static_cast<double>(N) * h * w * C * 2.0 * act_planes(plan->planes) * 5;
        snprintf(op.name, sizeof op.name, "upsample2 %dx%dx%d", h, w, C);
        plan->ops.push_back(op);
    }
    const ConvParam& P(const std::string& name) {
        auto it = ctx->conv_index.find(name);
        if (it == ctx->conv_index.end()) {
            rc = fail(ctx, FISR_E_INVALID, "unknown conv %s", name.c_str());
            static ConvParam dummy;
            return dummy;
        }
        return ctx->params[it->second];
    }

    // res_block (ops.py:39-44) x2 + trailing ReLU, given n0 (raw fp32) and relu(n0) (act): used by enc / dec levels.
    // Final activation relu(n2) goes to `dst` (channel offset dst_off of a dst_cs-channel buffer).
    ResRec two_res_blocks(const std::string& p, ActBuf a0, const float* n0, int c, int N, int H, int W, ActBuf dst,
                          int dst_cs, int dst_off, ActBuf pooled = ActBuf{}) {
        ActBuf a1 = act(N, H, W, c), a2 = act(N, H, W, c), a3 = act(N, H, W, c);
        float* n1 = f32(N, H, W, c);
        ConvOut o;
        o = ConvOut{}; o.act = a1; o.act_cs = c;
        conv(P(p + "/res_block/0/conv/0"), a0, c, 0, N, H, W, o, p + "/res_block/0/conv/0");
        o = ConvOut{}; o.res = n0; o.res_cs = c; o.raw = n1; o.raw_cs = c; o.act = a2; o.act_cs = c;
        conv(P(p + "/res_block/0/conv/1"), a1, c, 0, N, H, W, o, p + "/res_block/0/conv/1");
        o = ConvOut{}; o.act = a3; o.act_cs = c;
        conv(P(p + "/res_block/1/conv/0"), a2, c, 0, N, H, W, o, p + "/res_block/1/conv/0");
        o = ConvOut{}; o.res = n1; o.res_cs = c; o.act = dst; o.act_cs = dst_cs; o.act_off1 = dst_off;
        o.pool = pooled; o.pool_cs = c;              // encoder levels: max_pool(skip) written by the same epilogue (ops.py:52-54)
        conv(P(p + "/res_block/1/conv/1"), a3, c, 0, N, H, W, o, p + "/res_block/1/conv/1");
        return ResRec{a1, a2, a3};
    }

    // Enc_level_res (ops.py:48-55): skip = relu(...) lands in channels [c, 2c) of the decoder's concat buffer
    // (virtual tf.concat of ops.py:71); the 2x2 max-pooled copy feeds the next level down.
    ActBuf enc_level(const std::string& p, ActBuf x, int x_cs, int c, int N, int H, int W, ActBuf cat, EncRec* rec) {
        ActBuf a0 = act(N, H, W, c);
        float* n0 = f32(N, H, W, c);
        ConvOut o; o.raw = n0; o.raw_cs = c; o.act = a0; o.act_cs = c;
        conv(P(p + "/conv/0"), x, x_cs, 0, N, H, W, o, p + "/conv/0");
        ActBuf pooled = act(N, H / 2, W / 2, c);
        const ResRec rb = two_res_blocks(p, a0, n0, c, N, H, W, cat, 2 * c, c, pooled);      // max_pool(skip) rides in the last epilogue
        *rec = EncRec{p, x, x_cs, c, H, W, a0, cat, pooled, rb};
        return pooled;
    }

    // Dec_level_res (ops.py:67-76): x is [N, H/2, W/2, c1]; cat already holds the skip in channels [c, 2c).
    ActBuf dec_level(const std::string& p, ActBuf x, int c1, int c, int N, int H, int W, ActBuf cat, DecRec* rec) {
        ActBuf up = act(N, H, W, c1);
        upsample(x, up, N, H / 2, W / 2, c1);
        ConvOut o; o.act = cat; o.act_cs = 2 * c; o.act_off1 = 0;
        conv(P(p + "/resize"), up, c1, 0, N, H, W, o, p + "/resize");
        ActBuf a0 = act(N, H, W, c);
        float* n0 = f32(N, H, W, c);
        o = ConvOut{}; o.raw = n0; o.raw_cs = c; o.act = a0; o.act_cs = c;
        conv(P(p + "/conv/0"), cat, 2 * c, 0, N, H, W, o, p + "/conv/0");
        ActBuf out = act(N, H, W, c);
        const ResRec rb = two_res_blocks(p, a0, n0, c, N, H, W, out, c, 0);
        *rec = DecRec{p, x, up, cat, a0, out, rb, c1, c, H, W};
        return out;
    }

    // FI-SR / SR head (FISRnet.py:95-106).  pred is [N,2H,2W,9]; next (may be null) is the next level's input buffer.
    void head(const std::string& p, ActBuf x, int N, int H, int W, int cout, float* pred, ActBuf next, HeadRec* rec) {
        const int c = CH;
        ActBuf a0 = act(N, H, W, c), a1 = act(N, H, W, c), a2 = act(N, H, W, c);
        float* m0 = f32(N, H, W, c);
        ConvOut o; o.raw = m0; o.raw_cs = c; o.act = a0; o.act_cs = c;
        conv(P(p + "/conv/0"), x, c, 0, N, H, W, o, p + "/conv/0");
        o = ConvOut{}; o.act = a1; o.act_cs = c;
        conv(P(p + "/res_block/0/conv/0"), a0, c, 0, N, H, W, o, p + "/res_block/0/conv/0");
        o = ConvOut{}; o.res = m0; o.res_cs = c; o.act = a2; o.act_cs = c;
        conv(P(p + "/res_block/0/conv/1"), a1, c, 0, N, H, W, o, p + "/res_block/0/conv/1");
        if (plan->planes == 3) {
            // f16f8 (inference): conv/1 keeps its 256 channels at input resolution and conv/2 runs there too, with the
            // depth_to_space folded into its weights: 16 instead of 36 (tap, K block) products per output group, no 2R-resolution
            // activation tensor, and conv/1 stores plain NHWC rows.
            ActBuf t256 = act(N, H, W, 4 * c);
            o = ConvOut{}; o.act = t256; o.act_cs = 4 * c;             // relu commutes with depth_to_space (FISRnet.py:99,105)
            conv(P(p + "/conv/1"), a2, c, 0, N, H, W, o, p + "/conv/1");
            o = ConvOut{}; o.scalar = true; o.relu = false; o.ps = true;
            o.raw = pred; o.raw_cs = 12;                               // 12-float records: [FI-SR 0..2 -, SR 0..2 -, FI-SR 3..5 -]
            // (channels 29..37 of the next level's input are filled from pred by one pass after both heads, see level())
            if (cout == 6) { o.raw_off0 = 0; o.raw_off1 = 8; o.act_split = 3; o.act_off0 = IN_CH; o.act_off1 = IN_CH + 3; }
            else           { o.raw_off1 = 4; o.act_split = 0; o.act_off1 = IN_CH + 3; }
            conv_ps(P(p + "/conv/2"), t256, N, H, W, o, p + "/conv/2");
            *rec = HeadRec{p, x, a0, a1, a2, t256, cout};
            return;
        }
        ActBuf shuf = act(N, 2 * H, 2 * W, c);
        o = ConvOut{}; o.act = shuf; o.act_cs = c; o.d2s = true;       // relu + depth_to_space (FISRnet.py:99,105)
        conv(P(p + "/conv/1"), a2, c, 0, N, H, W, o, p + "/conv/1");
        // conv/2: FI-SR's 6 channels go to pred[..., 0:3] and [6:9], SR's 3 to [3:6] (FISRnet.py:107-108);
        // the same values feed channels 29.. of the next level's input (FISRnet.py:113,144), without ReLU.
        o = ConvOut{}; o.scalar = true; o.relu = false;
        o.raw = pred; o.raw_cs = 9;
        o.act = next; o.act_cs = 64;
        if (cout == 6) { o.raw_split = 3; o.raw_off0 = 0; o.raw_off1 = 3; o.act_split = 3; o.act_off0 = IN_CH; o.act_off1 = IN_CH + 3; }
        else           { o.raw_split = 0; o.raw_off1 = 3; o.act_split = 0; o.act_off1 = IN_CH + 3; }
        conv(P(p + "/conv/2"), shuf, c, 0, N, 2 * H, 2 * W, o, p + "/conv/2");
        *rec = HeadRec{p, x, a0, a1, a2, shuf, cout};
    }

    void level(int lvl, ActBuf in, int N, int H, int W, float* pred, ActBuf next) {
        const std::string p = "FISRnet/level_" + std::to_string(lvl);
        LevelRec& R = plan->rec[lvl - 1];
        R.lvl = lvl; R.N = N; R.H = H; R.W = W; R.in = in;
        ActBuf cat0 = act(N, H, W, 2 * CH), cat1 = act(N, H / 2, W / 2, 4 * CH), cat2 = act(N, H / 4, W / 4, 8 * CH);
        ActBuf n = enc_level(p + "/enc/level_0", in, 64, CH, N, H, W, cat0, &R.enc[0]);
        n = enc_level(p + "/enc/level_1", n, CH, 2 * CH, N, H / 2, W / 2, cat1, &R.enc[1]);
        n = enc_level(p + "/enc/level_2", n, 2 * CH, 4 * CH, N, H / 4, W / 4, cat2, &R.enc[2]);
        {   // Bottleneck_res (ops.py:59-63)
            const std::string b = p + "/bottleneck";
            const int c = 8 * CH, h = H / 8, w = W / 8;
            ActBuf a0 = act(N, h, w, c), a1 = act(N, h, w, c), out = act(N, h, w, c);
            float* n0 = f32(N, h, w, c);
            ConvOut o; o.raw = n0; o.raw_cs = c; o.act = a0; o.act_cs = c;
            conv(P(b + "/conv/0"), n, 4 * CH, 0, N, h, w, o, b + "/conv/0");
            o = ConvOut{}; o.act = a1; o.act_cs = c;
            conv(P(b + "/res_block/0/conv/0"), a0, c, 0, N, h, w, o, b + "/res_block/0/conv/0");
            o = ConvOut{}; o.res = n0; o.res_cs = c; o.act = out; o.act_cs = c;
            conv(P(b + "/res_block/0/conv/1"), a1, c, 0, N, h, w, o, b + "/res_block/0/conv/1");
            R.bott = BottRec{b, n, a0, a1, out, h, w};
            n = out;
        }
        n = dec_level(p + "/dec/level_2", n, 8 * CH, 4 * CH, N, H / 4, W / 4, cat2, &R.dec[2]);
        n = dec_level(p + "/dec/level_1", n, 4 * CH, 2 * CH, N, H / 2, W / 2, cat1, &R.dec[1]);
        n = dec_level(p + "/dec/level_0", n, 2 * CH, CH, N, H, W, cat0, &R.dec[0]);
        head(p + "/FI-SR", n, N, H, W, 6, pred, next, &R.head[0]);
        head(p + "/SR", n, N, H, W, 3, pred, next, &R.head[1]);
        if (plan->planes == 3 && next.p && rc == FISR_OK) {       // img_l{2,3} = concat(.., pred) (FISRnet.py:113,144)
            Op op{};
            op.kind = OP_PRED2NEXT; op.pred = pred; op.out = next; op.N = N; op.H = 2 * H; op.W = 2 * W;
            op.bytes = static_cast<double>(N) * 4 * H * W * (48 + 36);
            snprintf(op.name, sizeof op.name, "pred -> next level input %dx%d", 2 * H, 2 * W);
            plan->ops.push_back(op);
        }
    }
    // ================================================================ backward (see "Backward pass" in DESIGN.md)
    size_t max_partial = 0;

    ConvParam* PP(const std::string& name) {
        auto it = ctx->conv_index.find(name);
        if (it == ctx->conv_index.end()) { rc = fail(ctx, FISR_E_INVALID, "unknown conv %s", name.c_str()); return nullptr; }
        return &ctx->params[it->second];
    }

    // Data gradient of conv `name`: dx = conv3x3(dy, rot180(w)^T) through the forward kernel with the transposed operand
    // planes, gated by [mask > 0] and accumulated onto `res` in the epilogue (EPI_MASK / EPI_RES / EPI_S2D).
    void dconv(const std::string& name, ActBuf gin, int gin_cs, int gin_off, int N, int H, int W, ConvOut o) {
        ConvParam* p = PP(name);
        if (!p) return;
        o.relu = false;
        Op op{};
        if (!make_conv(bwd_view(*p), gin, gin_cs, gin_off, N, H, W, o, name + " [dgrad]", &op)) return;
        fisr_ctx* c = ctx;
        BwdOp b;
        b.name = op.name; b.flops = op.flops;
        b.run = [c, op](cudaStream_t st) -> int {
            const cudaError_t e = launch_conv3x3(op.conv, c->num_sms, st);
            return e == cudaSuccess ? FISR_OK : fail(c, FISR_E_CUDA, "dgrad launch failed: %s", cudaGetErrorString(e));
        };
        plan->bwd.push_back(std::move(b));
    }

    // Weight + bias gradient of conv `name` from its forward input x and the gradient dy of its output.
    void wgrad(const std::string& name, ActBuf x, int x_cs, int x_off, ActBuf dy, int dy_cs, int dy_off, int N, int H, int W) {
        ConvParam* p = PP(name);
        if (!p || rc != FISR_OK) return;
        WgradLaunch L{};
        plan_wgrad(N, H, W, p->KB, p->OBk, ctx->num_sms, ctx->wgrad_exact, &L);
        L.args.err = ctx->d_err; L.args.x_coff = x_off; L.args.dy_coff = dy_off;
        if (!encode_act(&L.tmX_hi, x.p, x_cs, N, H, W, kWgTW + 2, kWgTH + 2)) return;
        if (!encode_act(&L.tmX_lo, x.p + x.plane, x_cs, N, H, W, kWgTW + 2, kWgTH + 2)) return;
        if (!encode_act(&L.tmD_hi, dy.p, dy_cs, N, H, W, kWgTW, kWgTH)) return;
        if (!encode_act(&L.tmD_lo, dy.p + dy.plane, dy_cs, N, H, W, kWgTW, kWgTH)) return;
        const size_t npix = static_cast<size_t>(N) * H * W;
        max_partial = std::max(max_partial, L.partial_floats);
        fisr_ctx* c = ctx;
        Plan* pl = plan;
        BwdOp b;
        b.name = name + " [wgrad]";
        b.flops = 2.0 * 9 * p->cin * p->cout * static_cast<double>(npix);
        b.launches = 2;
        b.run = [c, pl, L, p](cudaStream_t st) -> int {
            WgradLaunch l = L;
            l.args.partial = pl->wg_partial;
            l.args.bias_partial = pl->wg_partial + l.bias_offset;
            const cudaError_t e = launch_wgrad3x3(l, st);
            if (e != cudaSuccess) return fail(c, FISR_E_CUDA, "wgrad launch failed: %s", cudaGetErrorString(e));
            launch_wgrad_reduce(l, pl->wg_partial, p->cin, p->cout, 1.f / pl->loss_scale, p->g_w, p->g_b, st);
            return FISR_OK;
        };
        plan->bwd.push_back(std::move(b));
    }

    void push_simple(const std::string& name, std::function<void(cudaStream_t)> fn) {
        BwdOp b;
        b.name = name;
        b.run = [fn](cudaStream_t st) -> int { fn(st); return FISR_OK; };
        plan->bwd.push_back(std::move(b));
    }

    // Backward of two res_blocks (ops.py:39-44 twice): given G(n2) planes g2 + fp32 r2, returns G(n0).
    //   n1 = n0 + conv1(relu(conv0(relu(n0)))),  n2 = n1 + conv1'(relu(conv0'(relu(n1))));  a0 = relu(n0), rb = {a1, a2, a3}
    ActBuf two_res_blocks_bwd(const std::string& p, ActBuf a0, const ResRec& rb, ActBuf g2, const float* r2, int c, int N, int H, int W) {
        ConvOut o;
        wgrad(p + "/res_block/1/conv/1", rb.a3, c, 0, g2, c, 0, N, H, W);
        ActBuf gc = act(N, H, W, c);
        o = ConvOut{}; o.act = gc; o.act_cs = c; o.mask = rb.a3; o.mask_cs = c;
        dconv(p + "/res_block/1/conv/1", g2, c, 0, N, H, W, o);
        wgrad(p + "/res_block/1/conv/0", rb.a2, c, 0, gc, c, 0, N, H, W);
        ActBuf g1 = act(N, H, W, c);
        float* r1 = f32(N, H, W, c);
        o = ConvOut{}; o.act = g1; o.act_cs = c; o.raw = r1; o.raw_cs = c; o.mask = rb.a2; o.mask_cs = c; o.res = r2; o.res_cs = c;
        dconv(p + "/res_block/1/conv/0", gc, c, 0, N, H, W, o);
        wgrad(p + "/res_block/0/conv/1", rb.a1, c, 0, g1, c, 0, N, H, W);
        ActBuf gc0 = act(N, H, W, c);
        o = ConvOut{}; o.act = gc0; o.act_cs = c; o.mask = rb.a1; o.mask_cs = c;
        dconv(p + "/res_block/0/conv/1", g1, c, 0, N, H, W, o);
        wgrad(p + "/res_block/0/conv/0", a0, c, 0, gc0, c, 0, N, H, W);
        ActBuf g0 = act(N, H, W, c);
        o = ConvOut{}; o.act = g0; o.act_cs = c; o.mask = a0; o.mask_cs = c; o.res = r1; o.res_cs = c;
        dconv(p + "/res_block/0/conv/0", gc0, c, 0, N, H, W, o);
        return g0;
    }

    // Backward of one level (FISRnet.py:84-108 and the same blocks of levels 2, 3).  gp = dy operand of the two conv/2
    // heads [N,2H,2W,128]; returns G(level input) [N,H,W,64] when want_input_grad (levels 2, 3: channels 29..37 carry the
    // gradient into the previous level's prediction).
    ActBuf level_bwd(const LevelRec& R, ActBuf gp, bool want_input_grad) {
        const int N = R.N, H = R.H, W = R.W, c = CH;
        ConvOut o;
        // ---- heads: out = conv2(d2s(relu(conv1(relu(m1))))), m1 = m0 + rb(m0), m0 = conv0(x)
        float* r_n = f32(N, H, W, c);
        ActBuf g_n = act(N, H, W, c);
        for (int hd = 0; hd < 2; ++hd) {
            const HeadRec& Hd = R.head[hd];
            const int off = hd == 0 ? 0 : 64;      // gp has 128 channels: FI-SR's dy in [0,6), SR's in [64,67)
            wgrad(Hd.p + "/conv/2", Hd.shuf, c, 0, gp, 2 * c, off, N, 2 * H, 2 * W);
            ActBuf g_c2 = act(N, H, W, 4 * c);
            o = ConvOut{}; o.act = g_c2; o.act_cs = 4 * c; o.mask = Hd.shuf; o.mask_cs = c; o.s2d = true;
            dconv(Hd.p + "/conv/2", gp, 2 * c, off, N, 2 * H, 2 * W, o);
            wgrad(Hd.p + "/conv/1", Hd.a2, c, 0, g_c2, 4 * c, 0, N, H, W);
            float* r_m1 = f32(N, H, W, c);
            ActBuf g_m1 = act(N, H, W, c);
            o = ConvOut{}; o.act = g_m1; o.act_cs = c; o.raw = r_m1; o.raw_cs = c; o.mask = Hd.a2; o.mask_cs = c;
            dconv(Hd.p + "/conv/1", g_c2, 4 * c, 0, N, H, W, o);
            wgrad(Hd.p + "/res_block/0/conv/1", Hd.a1, c, 0, g_m1, c, 0, N, H, W);
            ActBuf g_c0 = act(N, H, W, c);
            o = ConvOut{}; o.act = g_c0; o.act_cs = c; o.mask = Hd.a1; o.mask_cs = c;
            dconv(Hd.p + "/res_block/0/conv/1", g_m1, c, 0, N, H, W, o);
            wgrad(Hd.p + "/res_block/0/conv/0", Hd.a0, c, 0, g_c0, c, 0, N, H, W);
            ActBuf g_m0 = act(N, H, W, c);
            o = ConvOut{}; o.act = g_m0; o.act_cs = c; o.mask = Hd.a0; o.mask_cs = c; o.res = r_m1; o.res_cs = c;
            dconv(Hd.p + "/res_block/0/conv/0", g_c0, c, 0, N, H, W, o);
            wgrad(Hd.p + "/conv/0", Hd.x, c, 0, g_m0, c, 0, N, H, W);
            // both heads read the decoder output x = relu(n2): the second launch adds the first one's result in place
            o = ConvOut{}; o.act = g_n; o.act_cs = c; o.raw = r_n; o.raw_cs = c; o.mask = Hd.x; o.mask_cs = c;
            if (hd == 1) { o.res = r_n; o.res_cs = c; }
            dconv(Hd.p + "/conv/0", g_m0, c, 0, N, H, W, o);
        }
        // ---- decoder levels 0, 1, 2 (ops.py:67-76)
        ActBuf g_out = g_n;
        const float* r_out = r_n;
        ActBuf g_cat[3];
        for (int k = 0; k < 3; ++k) {
            const DecRec& D = R.dec[k];
            const int h = D.H, w = D.W, cc = D.c, c1 = D.c1;
            ActBuf g0 = two_res_blocks_bwd(D.p, D.a0, D.rb, g_out, r_out, cc, N, h, w);
            wgrad(D.p + "/conv/0", D.cat, 2 * cc, 0, g0, cc, 0, N, h, w);
            g_cat[k] = act(N, h, w, 2 * cc);
            o = ConvOut{}; o.act = g_cat[k]; o.act_cs = 2 * cc; o.mask = D.cat; o.mask_cs = 2 * cc;
            dconv(D.p + "/conv/0", g0, cc, 0, N, h, w, o);
            wgrad(D.p + "/resize", D.up, c1, 0, g_cat[k], 2 * cc, 0, N, h, w);
            ActBuf g_up = act(N, h, w, c1);
            o = ConvOut{}; o.act = g_up; o.act_cs = c1;
            dconv(D.p + "/resize", g_cat[k], 2 * cc, 0, N, h, w, o);
            ActBuf g_x = act(N, h / 2, w / 2, c1);
            float* r_x = f32(N, h / 2, w / 2, c1);
            {
                const ActBuf xin = D.x_in;
                push_simple("upsample2 backward " + std::to_string(h / 2) + "x" + std::to_string(w / 2) + "x" + std::to_string(c1),
                            [=](cudaStream_t st) { launch_upsample_bwd(g_up, xin, g_x, r_x, N, h / 2, w / 2, c1, st); });
            }
            g_out = g_x;
            r_out = r_x;
        }
        // ---- bottleneck (ops.py:59-63): out = relu(n1), n1 = n0 + conv1(relu(conv0(relu(n0)))), n0 = conv0(pooled)
        ActBuf g_pool;
        {
            const BottRec& Bt = R.bott;
            const int h = Bt.H, w = Bt.W, cc = 8 * c;
            wgrad(Bt.p + "/res_block/0/conv/1", Bt.a1, cc, 0, g_out, cc, 0, N, h, w);
            ActBuf gc = act(N, h, w, cc);
            o = ConvOut{}; o.act = gc; o.act_cs = cc; o.mask = Bt.a1; o.mask_cs = cc;
            dconv(Bt.p + "/res_block/0/conv/1", g_out, cc, 0, N, h, w, o);
            wgrad(Bt.p + "/res_block/0/conv/0", Bt.a0, cc, 0, gc, cc, 0, N, h, w);
            ActBuf g0 = act(N, h, w, cc);
            o = ConvOut{}; o.act = g0; o.act_cs = cc; o.mask = Bt.a0; o.mask_cs = cc; o.res = r_out; o.res_cs = cc;
            dconv(Bt.p + "/res_block/0/conv/0", gc, cc, 0, N, h, w, o);
            wgrad(Bt.p + "/conv/0", Bt.x, 4 * c, 0, g0, cc, 0, N, h, w);
            g_pool = act(N, h, w, 4 * c);
            o = ConvOut{}; o.act = g_pool; o.act_cs = 4 * c;
            dconv(Bt.p + "/conv/0", g0, cc, 0, N, h, w, o);
        }
        // ---- encoder levels 2, 1, 0 (ops.py:48-55)
        ActBuf g_in{};
        for (int k = 2; k >= 0; --k) {
            const EncRec& E = R.enc[k];
            const int h = E.H, w = E.W, cc = E.c;
            ActBuf g2 = act(N, h, w, cc);
            float* r2 = f32(N, h, w, cc);
            {
                const ActBuf skip = E.cat, gc = g_cat[k], gpl = g_pool;
                push_simple("maxpool2 backward " + std::to_string(h) + "x" + std::to_string(w) + "x" + std::to_string(cc),
                            [=](cudaStream_t st) { launch_pool_bwd(skip, gc, 2 * cc, cc, gpl, g2, r2, N, h, w, cc, st); });
            }
            ActBuf g0 = two_res_blocks_bwd(E.p, E.a0, E.rb, g2, r2, cc, N, h, w);
            wgrad(E.p + "/conv/0", E.x, E.x_cs, 0, g0, cc, 0, N, h, w);
            if (k > 0) {
                g_pool = act(N, h, w, cc / 2);
                o = ConvOut{}; o.act = g_pool; o.act_cs = cc / 2;
                dconv(E.p + "/conv/0", g0, cc, 0, N, h, w, o);
            } else if (want_input_grad) {
                g_in = act(N, h, w, 64);
                o = ConvOut{}; o.act = g_in; o.act_cs = 64;
                dconv(E.p + "/conv/0", g0, cc, 0, N, h, w, o);
            }
        }
        return g_in;
    }

    // Whole backward of one training step (4B passes batched as N): loss gradient per scale, level 3 -> 1.
    void backward(int B) {
        Plan* pl = plan;
        for (int l = 0; l < 3; ++l) pl->gp[l] = act(pl->N, 2 * pl->rec[l].H, 2 * pl->rec[l].W, 128, true);
        ActBuf g_next{};
        for (int l = 2; l >= 0; --l) {
            const LevelRec& R = pl->rec[l];
            const int hs = 2 * R.H, ws = 2 * R.W, stride = 4 >> l;
            const float wgt = static_cast<float>(4 >> l);
            const float* pred = pl->pred[l];
            const ActBuf extra = g_next, gp = pl->gp[l];
            const int LH = 2 * pl->H, LW = 2 * pl->W;
            push_simple("loss gradient level " + std::to_string(l + 1), [=](cudaStream_t st) {
                launch_loss_grad(pred, pl->label, extra, B, hs, ws, stride, LH, LW, wgt, pl->lam, pl->loss_scale, gp, st);
            });
            g_next = level_bwd(R, gp, l > 0);
        }
        pl->wg_partial = static_cast<float*>(alloc(max_partial * sizeof(float), false));
    }
};

int run_ops(fisr_ctx* ctx, Plan* plan, cudaStream_t st, std::vector<cudaEvent_t>* marks = nullptr) {
    size_t idx = 0;
    for (const Op& op : plan->ops) {
        if (marks) cudaEventRecord((*marks)[idx++], st);
        switch (op.kind) {
            case OP_CONV: {
                cudaError_t e = launch_conv3x3(op.conv, ctx->num_sms, st);
                if (e != cudaSuccess) return fail(ctx, FISR_E_CUDA, "conv launch failed: %s", cudaGetErrorString(e));
                break;
            }
            case OP_UPSAMPLE: launch_upsample2(op.in, op.out, op.N, op.H, op.W, op.C, plan->planes, st); break;
            case OP_PRED2NEXT: launch_pred_to_next(op.pred, op.out, static_cast<size_t>(op.N) * op.H * op.W, st); break;
        }
    }
    if (marks) cudaEventRecord((*marks)[idx], st);
    CUDA_TRY(ctx, cudaGetLastError());
    return FISR_OK;
}

int get_plan(fisr_ctx* ctx, int N, int H, int W, Plan** out) {
    if (N < 1 || H < 32 || W < 32 || H % 32 || W % 32)
        return fail(ctx, FISR_E_INVALID, "FISRnet.model needs N >= 1 and H, W multiples of 32 (got %d x %d x %d)", N, H, W);
    char key[64];
    snprintf(key, sizeof key, "%d_%d_%d_%d", N, H, W, ctx->planes);
    auto it = ctx->plans.find(key);
    if (it != ctx->plans.end()) { *out = it->second.get(); return FISR_OK; }

    for (auto& p : ctx->params) {
        int rc = ensure_packed(ctx, p, ctx->stream);
        if (rc != FISR_OK) return rc;
    }
    std::unique_ptr<Plan> plan(new Plan());
    plan->N = N; plan->H = H; plan->W = W; plan->planes = ctx->planes;
    Builder b{ctx, plan.get()};
    // level inputs: 64-channel buffers, channels >= 29 (38) stay zero (zero weights there, but 0 * NaN != 0)
    plan->in_lvl[0] = b.act(N, H / 4, W / 4, 64, true);
    plan->in_lvl[1] = b.act(N, H / 2, W / 2, 64, true);
    plan->in_lvl[2] = b.act(N, H, W, 64, true);
    plan->pred_cs = ctx->planes == 3 ? 12 : 9;
    plan->pred[0] = b.f32(N, H / 2, W / 2, plan->pred_cs);
    plan->pred[1] = b.f32(N, H, W, plan->pred_cs);
    plan->pred[2] = b.f32(N, 2 * H, 2 * W, plan->pred_cs);
    b.level(1, plan->in_lvl[0], N, H / 4, W / 4, plan->pred[0], plan->in_lvl[1]);
    b.level(2, plan->in_lvl[1], N, H / 2, W / 2, plan->pred[1], plan->in_lvl[2]);
    b.level(3, plan->in_lvl[2], N, H, W, plan->pred[2], ActBuf{});
    if (b.rc != FISR_OK) return b.rc;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));

    if (ctx->use_graph) {
        cudaGraph_t g = nullptr;
        CUDA_TRY(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        int rc = run_ops(ctx, plan.get(), ctx->stream);
        cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
        if (rc != FISR_OK) { if (g) cudaGraphDestroy(g); return rc; }
        if (e != cudaSuccess) return fail(ctx, FISR_E_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
        e = cudaGraphInstantiate(&plan->graph, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) return fail(ctx, FISR_E_CUDA, "graph instantiate failed: %s", cudaGetErrorString(e));
    }
    *out = plan.get();
    ctx->plans[key] = std::move(plan);
    return FISR_OK;
}

// Runs the conv stack of `plan` on stream st (inputs already packed into plan->in_lvl).
int run_plan(fisr_ctx* ctx, Plan* plan, cudaStream_t st) {
    ctx->last_plan = plan;
    ctx->launches += static_cast<long long>(plan->ops.size());
    if (plan->graph) {
        // an instantiated graph may be launched into any stream, not only the one it was captured on
        CUDA_TRY(ctx, cudaGraphLaunch(plan->graph, st));
        return FISR_OK;
    }
    return run_ops(ctx, plan, st);
}

// Host restatement of utils.get_HW_boundary / trim_patch_boundary (utils.py:118-159) for the whole tile grid.
struct TileGeom {
    int h, w, sH, sW;
    struct T { int ylo, yhi, xlo, xhi, trim_y, trim_x; };
    std::vector<T> tiles;
};
TileGeom tile_geometry(int H, int W, int pH, int pW) {
    const int pb = 32;
    TileGeom g;
    g.h = H - H % (32 * pH);                       // FISRnet.py:1006-1007
    g.w = W - W % (32 * pW);
    g.sH = g.h / pH;
    g.sW = g.w / pW;
    for (int p = 0; p < pH * pW; ++p) {
        const int iy = p / pW, ix = p % pW;         // FISRnet.py:1029-1030
        TileGeom::T t;
        t.ylo = std::max(iy * g.sH - pb, 0);
        t.yhi = std::min((iy + 1) * g.sH + pb, g.h);
        t.xlo = std::max(ix * g.sW - pb, 0);
        t.xhi = std::min((ix + 1) * g.sW + pb, g.w);
        t.trim_y = (iy * g.sH < pb) ? 0 : pb * 2;   // utils.py:142-145 (sf = 2)
        t.trim_x = (ix * g.sW < pb) ? 0 : pb * 2;   // utils.py:150-153
        g.tiles.push_back(t);
    }
    return g;
}

// Runs a list of (window, tile) units: unit id = window * (pH*pW) + tile.  Units with equal tile size are batched
// into one forward.  layout 0: canvas is [B, 2h, 2w, 9], every tile pasted at its place in its window's frame
// (FISRnet.py:1056-1057); layout 1: canvas is [n_units, 2sH, 2sW, 9], unit i of the list in slot i (the send buffer
// of the multi-GPU all-gather).
int units_impl(fisr_ctx* ctx, const uint8_t* d_frames, const float* d_flow, const float* d_warp, int B, int H, int W,
               int pH, int pW, const int* units, int n_units, int layout, uint8_t* d_canvas_u8, float* d_canvas_f32,
               cudaStream_t st) {
    if (B < 1 || pH < 1 || pW < 1 || H < 32 * pH || W < 32 * pW)
        return fail(ctx, FISR_E_INVALID, "bad window batch %d or tile grid %dx%d for %dx%d", B, pH, pW, H, W);
    if (n_units < 0 || (n_units > 0 && !units)) return fail(ctx, FISR_E_INVALID, "bad unit list");
    const int T = pH * pW;
    for (int i = 0; i < n_units; ++i)
        if (units[i] < 0 || units[i] >= B * T) return fail(ctx, FISR_E_INVALID, "unit %d outside %d windows x %d tiles", units[i], B, T);
    const TileGeom g = tile_geometry(H, W, pH, pW);
    const int OH = layout == 0 ? 2 * g.h : 2 * g.sH, OW = layout == 0 ? 2 * g.w : 2 * g.sW;
    std::vector<bool> done(n_units, false);
    for (int i = 0; i < n_units; ++i) {
        if (done[i]) continue;
        const TileGeom::T& ti = g.tiles[units[i] % T];
        const int th = ti.yhi - ti.ylo, tw = ti.xhi - ti.xlo;
        TileList tl{};
        for (int j = i; j < n_units && tl.count < kMaxTiles; ++j) {
            if (done[j]) continue;
            const int q = units[j] % T;
            const TileGeom::T& tj = g.tiles[q];
            if (tj.yhi - tj.ylo != th || tj.xhi - tj.xlo != tw) continue;
            const int k = tl.count++;
            tl.win[k] = units[j] / T;
            tl.ylo[k] = tj.ylo; tl.xlo[k] = tj.xlo;
            tl.trim_y[k] = tj.trim_y; tl.trim_x[k] = tj.trim_x;
            if (layout == 0) { tl.out_img[k] = units[j] / T; tl.out_y[k] = (q / pW) * g.sH * 2; tl.out_x[k] = (q % pW) * g.sW * 2; }
            else             { tl.out_img[k] = j; tl.out_y[k] = 0; tl.out_x[k] = 0; }
            done[j] = true;
        }
        Plan* plan = nullptr;
        int rc = get_plan(ctx, tl.count, th, tw, &plan);
        if (rc != FISR_OK) return rc;
        // frames are H x W on the device (uncropped); the crop is a view (img[:h, :w], FISRnet.py:1008)
        launch_tile_pack(d_frames, d_flow, d_warp, H, W, tl, th, tw, ctx->d_lut255, plan->in_lvl[2], plan->in_lvl[1],
                         plan->in_lvl[0], plan->planes, st);
        ctx->launches++;
        rc = run_plan(ctx, plan, st);
        if (rc != FISR_OK) return rc;
        if (d_canvas_u8)
            launch_tile_unpack_u8(plan->pred[2], plan->pred_cs, tl, 2 * th, 2 * tw, d_canvas_u8, OH, OW, 2 * g.sH, 2 * g.sW, st);
        if (d_canvas_f32)
            launch_tile_unpack_f32(plan->pred[2], plan->pred_cs, tl, 2 * th, 2 * tw, d_canvas_f32, OH, OW, 2 * g.sH, 2 * g.sW, st);
        ctx->launches++;
        CUDA_TRY(ctx, cudaGetLastError());
    }
    return FISR_OK;
}

int window_impl(fisr_ctx* ctx, const uint8_t* d_frames, const float* d_flow, const float* d_warp, int H, int W, int pH,
                int pW, int tile_first, int tile_count, uint8_t* d_canvas_u8, float* d_canvas_f32, cudaStream_t st) {
    if (pH < 1 || pW < 1) return fail(ctx, FISR_E_INVALID, "bad tile grid %dx%d", pH, pW);
    if (tile_first < 0 || tile_count < 0 || tile_first + tile_count > pH * pW)
        return fail(ctx, FISR_E_INVALID, "tile range [%d,+%d) outside the %dx%d grid", tile_first, tile_count, pH, pW);
    std::vector<int> units(tile_count);
    for (int i = 0; i < tile_count; ++i) units[i] = tile_first + i;
    return units_impl(ctx, d_frames, d_flow, d_warp, 1, H, W, pH, pW, units.data(), tile_count, 0, d_canvas_u8, d_canvas_f32, st);
}


// ---------------------------------------------------------------- training plumbing
int ensure_train_params(fisr_ctx* ctx, cudaStream_t st) {
    for (auto& p : ctx->params) {
        if (!p.d_wpT) {
            CUDA_TRY(ctx, cudaMalloc(&p.d_wpT, static_cast<size_t>(2) * p.OBk * 9 * p.cin_pad * 64 * sizeof(__half)));
            CUDA_TRY(ctx, cudaMalloc(&p.g_w, static_cast<size_t>(9) * p.cin * p.cout * 4));
            CUDA_TRY(ctx, cudaMalloc(&p.g_b, static_cast<size_t>(p.cout) * 4));
            p.packedT = false;
        }
        if (!p.packedT) {
            launch_prep_weights_dgrad(p.d_w, p.d_wpT, p.cin, p.cout, p.OBk, p.cin_pad, st);
            ctx->launches++;
            p.packedT = true;
        }
    }
    CUDA_TRY(ctx, cudaGetLastError());
    return FISR_OK;
}

int ensure_backward(fisr_ctx* ctx, Plan* plan, int B) {
    if (plan->has_bwd) return FISR_OK;
    if (plan->planes != 2) return fail(ctx, FISR_E_INVALID, "training needs precision f16x3 (split operands)");
    Builder b{ctx, plan};
    b.backward(B);
    if (b.rc != FISR_OK) return b.rc;
    plan->B = B;
    // Loss scale: a power of two near the element count of one finest-scale frame stack, so that d loss / d pred is
    // O(1) in the fp16 (hi, lo) gradient planes; divided out again by the wgrad / bias-grad reductions.
    const double n1 = static_cast<double>(B) * (2.0 * plan->H) * (2.0 * plan->W) * 3.0;
    plan->loss_scale = ctx->loss_scale_override > 0.f ? ctx->loss_scale_override : std::exp2(std::floor(std::log2(n1)));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    plan->has_bwd = true;
    return FISR_OK;
}

int run_backward(fisr_ctx* ctx, Plan* plan, cudaStream_t st) {
    static const bool debug_sync = getenv("FISR_DEBUG_SYNC") && getenv("FISR_DEBUG_SYNC")[0] == '1';
    for (const BwdOp& op : plan->bwd) {
        const int rc = op.run(st);
        if (rc != FISR_OK) return rc;
        ctx->launches += op.launches;
        if (debug_sync) {     // bring-up aid: attribute an asynchronous fault to the op that caused it
            const cudaError_t e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) return fail(ctx, FISR_E_CUDA, "backward op '%s' failed: %s", op.name.c_str(), cudaGetErrorString(e));
        }
    }
    CUDA_TRY(ctx, cudaGetLastError());
    return FISR_OK;
}

int ensure_adam_slots(fisr_ctx* ctx, cudaStream_t st) {
    for (auto& p : ctx->params) {
        if (p.m_w) continue;
        const size_t wn = static_cast<size_t>(9) * p.cin * p.cout, bn = p.cout;
        CUDA_TRY(ctx, cudaMalloc(&p.m_w, wn * 4)); CUDA_TRY(ctx, cudaMalloc(&p.v_w, wn * 4));
        CUDA_TRY(ctx, cudaMalloc(&p.m_b, bn * 4)); CUDA_TRY(ctx, cudaMalloc(&p.v_b, bn * 4));
        CUDA_TRY(ctx, cudaMemsetAsync(p.m_w, 0, wn * 4, st)); CUDA_TRY(ctx, cudaMemsetAsync(p.v_w, 0, wn * 4, st));
        CUDA_TRY(ctx, cudaMemsetAsync(p.m_b, 0, bn * 4, st)); CUDA_TRY(ctx, cudaMemsetAsync(p.v_b, 0, bn * 4, st));
    }
    return FISR_OK;
}

// Device table over all 276 tensors for the multi-tensor kernels (rebuilt per call: the gradient pointers may be the caller's).
int upload_mt_table(fisr_ctx* ctx, const std::vector<const float*>& grads, cudaStream_t st, unsigned* blocks) {
    std::vector<MtTensor>& t = ctx->mt_host;
    t.clear();
    unsigned blk = 0;
    for (size_t i = 0; i < ctx->params.size(); ++i) {
        ConvParam& p = ctx->params[i];
        const unsigned long long wn = static_cast<unsigned long long>(9) * p.cin * p.cout, bn = p.cout;
        t.push_back(MtTensor{p.d_w, grads[2 * i], p.m_w, p.v_w, wn, blk});
        blk += static_cast<unsigned>((wn + kMtChunk - 1) / kMtChunk);
        t.push_back(MtTensor{p.d_b, grads[2 * i + 1], p.m_b, p.v_b, bn, blk});
        blk += static_cast<unsigned>((bn + kMtChunk - 1) / kMtChunk);
    }
    if (!ctx->d_mt) CUDA_TRY(ctx, cudaMalloc(&ctx->d_mt, t.size() * sizeof(MtTensor)));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_mt, t.data(), t.size() * sizeof(MtTensor), cudaMemcpyHostToDevice, st));
    *blocks = blk;
    return FISR_OK;
}

// Operand planes follow the fp32 master copies: one multi-tensor launch for the split-mode forward planes and (when the
// training planes exist) the rotated-transposed dgrad planes; other precision modes re-pack conv by conv.
int repack_all(fisr_ctx* ctx, cudaStream_t st) {
    if (ctx->planes != 2) {
        for (auto& p : ctx->params) {
            p.packed = false; p.packedT = false;
            const int rc = ensure_packed(ctx, p, st);
            if (rc != FISR_OK) return rc;
        }
        return FISR_OK;
    }
    std::vector<MtPack> t;
    unsigned blk = 0;
    for (auto& p : ctx->params) {
        const unsigned long long pf = static_cast<unsigned long long>(p.KB) * 9 * p.cout_pad * 64;
        t.push_back(MtPack{p.d_w, p.d_wp, p.cin, p.cout, p.cout_pad, 0, pf, blk});
        blk += static_cast<unsigned>((pf + kMtChunk - 1) / kMtChunk);
        p.packed = true;
        if (p.d_wpT) {
            const unsigned long long pb = static_cast<unsigned long long>(p.OBk) * 9 * p.cin_pad * 64;
            t.push_back(MtPack{p.d_w, p.d_wpT, p.cin, p.cout, p.cin_pad, 1, pb, blk});
            blk += static_cast<unsigned>((pb + kMtChunk - 1) / kMtChunk);
            p.packedT = true;
        } else {
            p.packedT = false;
        }
    }
    if (ctx->mt_pack_cap < t.size()) {
        if (ctx->d_mt_pack) cudaFree(ctx->d_mt_pack);
        ctx->d_mt_pack = nullptr; ctx->mt_pack_cap = 0;
        CUDA_TRY(ctx, cudaMalloc(&ctx->d_mt_pack, t.size() * sizeof(MtPack)));
        ctx->mt_pack_cap = t.size();
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_mt_pack, t.data(), t.size() * sizeof(MtPack), cudaMemcpyHostToDevice, st));
    launch_mt_repack(ctx->d_mt_pack, static_cast<int>(t.size()), blk, st);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return FISR_OK;
}

// tf.train.AdamOptimizer update of all 276 tensors (FISRnet.py:489-491) + operand re-pack, asynchronous on stream st:
// 2 kernel launches (multi-tensor Adam, multi-tensor re-pack).
int adam_impl(fisr_ctx* ctx, const std::vector<const float*>& grads, float lr, float beta1, float beta2, float eps, cudaStream_t st) {
    int rc = ensure_adam_slots(ctx, st);
    if (rc != FISR_OK) return rc;
    const long long t = ++ctx->adam_t;
    const float lr_t = static_cast<float>(lr * std::sqrt(1.0 - std::pow(static_cast<double>(beta2), static_cast<double>(t))) /
                                          (1.0 - std::pow(static_cast<double>(beta1), static_cast<double>(t))));
    unsigned blocks = 0;
    if ((rc = upload_mt_table(ctx, grads, st, &blocks)) != FISR_OK) return rc;
    launch_mt_adam(ctx->d_mt, static_cast<int>(ctx->mt_host.size()), blocks, lr_t, beta1, beta2, eps, st);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return repack_all(ctx, st);
}

}  // namespace

// ================================================================ C ABI
extern "C" {

int fisr_create(int device, fisr_ctx** out) {
    if (!out) return fail(nullptr, FISR_E_INVALID, "fisr_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, FISR_E_CUDA, "no CUDA device: %s (fisr_b200 has no CPU path)", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, FISR_E_INVALID, "device %d out of range (%d devices)", device, ndev);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return fail(nullptr, FISR_E_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, FISR_E_CUDA, "fisr_b200 is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
    std::unique_ptr<fisr_ctx> ctx(new fisr_ctx());
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    Guard guard(device);
    cudaFree(0);
    cudaDriverEntryPointQueryResult q;
    void* fn = nullptr;
    if ((e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q)) != cudaSuccess || !fn)
        return fail(nullptr, FISR_E_CUDA, "cuTensorMapEncodeTiled unavailable: %s", cudaGetErrorString(e));
    ctx->encode = reinterpret_cast<EncodeTiledFn>(fn);
    if ((e = conv3x3_init()) != cudaSuccess)
        return fail(nullptr, FISR_E_CUDA, "conv kernel attribute setup failed: %s", cudaGetErrorString(e));
    if ((e = wgrad3x3_init()) != cudaSuccess)
        return fail(nullptr, FISR_E_CUDA, "wgrad kernel attribute setup failed: %s", cudaGetErrorString(e));
    const char* g = getenv("FISR_NO_GRAPH");
    ctx->use_graph = !(g && g[0] == '1');
    fisr_ctx* c = ctx.get();
    CUDA_TRY(nullptr, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUDA_TRY(nullptr, cudaMalloc(&c->d_err, sizeof(int)));
    CUDA_TRY(nullptr, cudaMemset(c->d_err, 0, sizeof(int)));
    {
        float lut[256];
        for (int i = 0; i < 256; ++i) lut[i] = static_cast<float>(static_cast<double>(i) / 255.0);   // FISRnet.py:1011
        CUDA_TRY(nullptr, cudaMalloc(&c->d_lut255, sizeof lut));
        CUDA_TRY(nullptr, cudaMemcpy(c->d_lut255, lut, sizeof lut, cudaMemcpyHostToDevice));
    }
    CUDA_TRY(nullptr, cudaMalloc(&c->d_zero_bias, 512 * sizeof(float)));
    CUDA_TRY(nullptr, cudaMemset(c->d_zero_bias, 0, 512 * sizeof(float)));
    CUDA_TRY(nullptr, cudaMalloc(&c->d_gmax, sizeof(unsigned)));
    const auto& inv = inventory();
    c->params.resize(inv.size());
    for (size_t i = 0; i < inv.size(); ++i) {
        ConvParam& p = c->params[i];
        p.cin = inv[i].cin; p.cout = inv[i].cout;
        p.KB = (p.cin + 63) / 64;
        p.cout_pad = p.cout <= 16 ? 16 : (p.cout + 63) / 64 * 64;
        p.cin_pad = p.KB * 64;
        p.OBk = (p.cout + 63) / 64;
        p.head2 = p.cin == 64 && p.cout <= 16;                 // FI-SR/conv/2 (64 -> 6) and SR/conv/2 (64 -> 3)
        p.ps_cout_pad = 4 * p.cout <= 16 ? 16 : 32;
        const size_t wn = static_cast<size_t>(9) * p.cin * p.cout;
        CUDA_TRY(nullptr, cudaMalloc(&p.d_w, wn * 4));
        CUDA_TRY(nullptr, cudaMemset(p.d_w, 0, wn * 4));
        CUDA_TRY(nullptr, cudaMalloc(&p.d_b, p.cout_pad * 4));
        CUDA_TRY(nullptr, cudaMemset(p.d_b, 0, p.cout_pad * 4));
        CUDA_TRY(nullptr, cudaMalloc(&p.d_wp, static_cast<size_t>(2) * p.KB * 9 * p.cout_pad * 64 * 2));
        c->conv_index[inv[i].name] = static_cast<int>(i);
    }
    CUDA_TRY(nullptr, cudaDeviceSynchronize());   // legacy-stream memsets above vs. the non-blocking context stream
    *out = ctx.release();
    return FISR_OK;
}

void fisr_destroy(fisr_ctx* ctx) {
    if (!ctx) return;
    Guard guard(ctx->device);
    cudaDeviceSynchronize();
    ctx->plans.clear();
    for (auto& p : ctx->params) {
        cudaFree(p.d_w); cudaFree(p.d_b); cudaFree(p.d_wp); cudaFree(p.d_wps); cudaFree(p.d_bps); cudaFree(p.d_wpps); cudaFree(p.m_w); cudaFree(p.v_w); cudaFree(p.m_b); cudaFree(p.v_b);
        cudaFree(p.d_wpT); cudaFree(p.g_w); cudaFree(p.g_b);
    }
    cudaFree(ctx->d_scalars);
    cudaFree(ctx->d_zero_bias);
    cudaFree(ctx->d_gmax);
    cudaFree(ctx->d_mt);
    cudaFree(ctx->d_mt_pack);
    for (void* s : ctx->stage) if (s) cudaFree(s);
    for (auto& sl : ctx->slots) {
        for (void* b : sl.in) if (b) cudaFree(b);
        if (sl.out) cudaFree(sl.out);
        if (sl.h_err) cudaFreeHost(sl.h_err);
        if (sl.h2d_done) cudaEventDestroy(sl.h2d_done);
        if (sl.compute_done) cudaEventDestroy(sl.compute_done);
        if (sl.d2h_done) cudaEventDestroy(sl.d2h_done);
    }
    if (ctx->h2d_stream) cudaStreamDestroy(ctx->h2d_stream);
    if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
    cudaFree(ctx->d_err);
    cudaFree(ctx->d_lut255);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* fisr_last_error(const fisr_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int fisr_set_precision(fisr_ctx* ctx, int precision) {
    if (!ctx) return FISR_E_INVALID;
    if (precision != FISR_PREC_F16X3 && precision != FISR_PREC_F16 && precision != FISR_PREC_F16F8)
        return fail(ctx, FISR_E_INVALID, "unknown precision %d", precision);
    const int planes = precision == FISR_PREC_F16X3 ? 2 : precision == FISR_PREC_F16F8 ? 3 : 1;   // kernel PLANES parameter
    if (planes != ctx->planes) {
        Guard guard(ctx->device);
        cudaDeviceSynchronize();
        ctx->plans.clear();
        ctx->last_plan = nullptr;
        ctx->planes = planes;
        for (auto& p : ctx->params) p.packed = false;
    }
    return FISR_OK;
}
int fisr_get_precision(const fisr_ctx* ctx) {
    return ctx && ctx->planes == 1 ? FISR_PREC_F16 : ctx && ctx->planes == 3 ? FISR_PREC_F16F8 : FISR_PREC_F16X3;
}

int fisr_num_params(void) { return static_cast<int>(param_names().size()); }
const char* fisr_param_name(int index) {
    const auto& n = param_names();
    return (index >= 0 && index < (int)n.size()) ? n[index].c_str() : nullptr;
}
int fisr_param_shape(int index, int dims[4]) {
    const auto& inv = inventory();
    if (index < 0 || index >= 2 * (int)inv.size() || !dims) return FISR_E_INVALID;
    const ParamDef& d = inv[index / 2];
    if (index % 2 == 0) { dims[0] = 3; dims[1] = 3; dims[2] = d.cin; dims[3] = d.cout; return 4; }
    dims[0] = d.cout; dims[1] = dims[2] = dims[3] = 1;
    return 1;
}

static int split_param_name(fisr_ctx* ctx, const char* name, int* conv, bool* is_w) {
    if (!name) return fail(ctx, FISR_E_INVALID, "parameter name is NULL");
    std::string s(name);
    if (s.size() < 3 || (s.substr(s.size() - 2) != "/w" && s.substr(s.size() - 2) != "/b"))
        return fail(ctx, FISR_E_INVALID, "parameter name must end in /w or /b: %s", name);
    auto it = ctx->conv_index.find(s.substr(0, s.size() - 2));
    if (it == ctx->conv_index.end()) return fail(ctx, FISR_E_INVALID, "unknown parameter %s", name);
    *conv = it->second;
    *is_w = s.back() == 'w';
    return FISR_OK;
}

int fisr_set_param(fisr_ctx* ctx, const char* name, const float* h_data, size_t count) {
    if (!ctx || !h_data) return FISR_E_INVALID;
    int ci; bool is_w;
    int rc = split_param_name(ctx, name, &ci, &is_w);
    if (rc != FISR_OK) return rc;
    Guard guard(ctx->device);
    ConvParam& p = ctx->params[ci];
    const size_t expect = is_w ? static_cast<size_t>(9) * p.cin * p.cout : static_cast<size_t>(p.cout);
    if (count != expect) return fail(ctx, FISR_E_INVALID, "%s has %zu elements, got %zu", name, expect, count);
    // Copy and re-pack on ONE stream: a pageable cudaMemcpy on the legacy stream may return before its DMA has
    // landed, and the context stream is non-blocking, so the pack kernel would race with it.
    CUDA_TRY(ctx, cudaDeviceSynchronize());
    CUDA_TRY(ctx, cudaMemcpyAsync(is_w ? p.d_w : p.d_b, h_data, count * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (is_w || p.head2) {
        p.packed = false;
        if (is_w) p.packedT = false;
        // plans hold pointers to the packed planes, which are rewritten in place: re-pack now
        rc = ensure_packed(ctx, p, ctx->stream);
        if (rc != FISR_OK) return rc;
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return FISR_OK;
}

int fisr_get_param(fisr_ctx* ctx, const char* name, float* h_data, size_t count) {
    if (!ctx || !h_data) return FISR_E_INVALID;
    int ci; bool is_w;
    int rc = split_param_name(ctx, name, &ci, &is_w);
    if (rc != FISR_OK) return rc;
    Guard guard(ctx->device);
    ConvParam& p = ctx->params[ci];
    const size_t expect = is_w ? static_cast<size_t>(9) * p.cin * p.cout : static_cast<size_t>(p.cout);
    if (count != expect) return fail(ctx, FISR_E_INVALID, "%s has %zu elements, got %zu", name, expect, count);
    CUDA_TRY(ctx, cudaMemcpyAsync(h_data, is_w ? p.d_w : p.d_b, count * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return FISR_OK;
}

// FISRnet.model's three outputs as dense [.., 9] tensors: the plan's buffers themselves, or compacted copies of the 12-float records.
int compact_preds(fisr_ctx* ctx, Plan* plan, const float* src[3], cudaStream_t st) {
    const size_t npx[3] = {static_cast<size_t>(plan->N) * (plan->H / 2) * (plan->W / 2), static_cast<size_t>(plan->N) * plan->H * plan->W,
                           static_cast<size_t>(plan->N) * plan->H * plan->W * 4};
    for (int l = 0; l < 3; ++l) {
        src[l] = plan->pred[l];
        if (plan->pred_cs == 9) continue;
        if (!plan->pred9[l]) {
            void* p = nullptr;
            CUDA_TRY(ctx, cudaMalloc(&p, npx[l] * 9 * sizeof(float)));
            plan->allocs.push_back(p);
            plan->pred9[l] = static_cast<float*>(p);
        }
        launch_pred_compact(plan->pred[l], plan->pred9[l], npx[l], st);
        ctx->launches++;
        src[l] = plan->pred9[l];
    }
    CUDA_TRY(ctx, cudaGetLastError());
    return FISR_OK;
}

int fisr_forward(fisr_ctx* ctx, const float* d_img, int N, int H, int W, float* d_l1, float* d_l2, float* d_l3,
                 void* stream) {
    if (!ctx || !d_img) return FISR_E_INVALID;
    Guard guard(ctx->device);
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    Plan* plan = nullptr;
    int rc = get_plan(ctx, N, H, W, &plan);
    if (rc != FISR_OK) return rc;
    launch_pack_input(d_img, N, H, W, IN_CH, plan->in_lvl[2], plan->in_lvl[1], plan->in_lvl[0], plan->planes, st);
    ctx->launches++;
    rc = run_plan(ctx, plan, st);
    if (rc != FISR_OK) return rc;
    const size_t n1 = static_cast<size_t>(N) * (H / 2) * (W / 2) * 9, n2 = static_cast<size_t>(N) * H * W * 9, n3 = n2 * 4;
    const float* src[3];
    if ((rc = compact_preds(ctx, plan, src, st)) != FISR_OK) return rc;
    if (d_l1) CUDA_TRY(ctx, cudaMemcpyAsync(d_l1, src[0], n1 * 4, cudaMemcpyDeviceToDevice, st));
    if (d_l2) CUDA_TRY(ctx, cudaMemcpyAsync(d_l2, src[1], n2 * 4, cudaMemcpyDeviceToDevice, st));
    if (d_l3) CUDA_TRY(ctx, cudaMemcpyAsync(d_l3, src[2], n3 * 4, cudaMemcpyDeviceToDevice, st));
    return FISR_OK;
}

int fisr_forward_host(fisr_ctx* ctx, const float* h_img, int N, int H, int W, float* h_l1, float* h_l2, float* h_l3) {
    if (!ctx || !h_img) return FISR_E_INVALID;
    Guard guard(ctx->device);
    Plan* plan = nullptr;
    int rc = get_plan(ctx, N, H, W, &plan);
    if (rc != FISR_OK) return rc;
    const size_t in_bytes = static_cast<size_t>(N) * H * W * IN_CH * 4;
    if ((rc = ensure_stage(ctx, 0, in_bytes)) != FISR_OK) return rc;
    cudaStream_t st = ctx->stream;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->stage[0], h_img, in_bytes, cudaMemcpyHostToDevice, st));
    launch_pack_input(static_cast<const float*>(ctx->stage[0]), N, H, W, IN_CH, plan->in_lvl[2], plan->in_lvl[1],
                      plan->in_lvl[0], plan->planes, st);
    ctx->launches++;
    rc = run_plan(ctx, plan, st);
    if (rc != FISR_OK) return rc;
    const size_t n1 = static_cast<size_t>(N) * (H / 2) * (W / 2) * 9, n2 = static_cast<size_t>(N) * H * W * 9, n3 = n2 * 4;
    const float* src[3];
    if ((rc = compact_preds(ctx, plan, src, st)) != FISR_OK) return rc;
    if (h_l1) CUDA_TRY(ctx, cudaMemcpyAsync(h_l1, src[0], n1 * 4, cudaMemcpyDeviceToHost, st));
    if (h_l2) CUDA_TRY(ctx, cudaMemcpyAsync(h_l2, src[1], n2 * 4, cudaMemcpyDeviceToHost, st));
    if (h_l3) CUDA_TRY(ctx, cudaMemcpyAsync(h_l3, src[2], n3 * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return check_kernel_error(ctx);
}

int fisr_window_device(fisr_ctx* ctx, const uint8_t* d_frames, const float* d_flow, const float* d_warp, int H, int W,
                       int pH, int pW, int tile_first, int tile_count, uint8_t* d_canvas, void* stream) {
    if (!ctx || !d_frames || !d_flow || !d_warp || !d_canvas) return FISR_E_INVALID;
    Guard guard(ctx->device);
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    return window_impl(ctx, d_frames, d_flow, d_warp, H, W, pH, pW, tile_first, tile_count, d_canvas, nullptr, st);
}

int fisr_units_device(fisr_ctx* ctx, const uint8_t* d_frames, const float* d_flow, const float* d_warp, int B, int H, int W,
                      int pH, int pW, const int* h_units, int n_units, int layout, uint8_t* d_out, void* stream) {
    if (!ctx || !d_frames || !d_flow || !d_warp || !d_out) return FISR_E_INVALID;
    if (layout != 0 && layout != 1) return fail(ctx, FISR_E_INVALID, "layout must be 0 (frames) or 1 (unit-major)");
    Guard guard(ctx->device);
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    return units_impl(ctx, d_frames, d_flow, d_warp, B, H, W, pH, pW, h_units, n_units, layout, d_out, nullptr, st);
}

int fisr_window_device_f32(fisr_ctx* ctx, const uint8_t* d_frames, const float* d_flow, const float* d_warp, int H,
                           int W, int pH, int pW, float* d_canvas, void* stream) {
    if (!ctx || !d_frames || !d_flow || !d_warp || !d_canvas) return FISR_E_INVALID;
    Guard guard(ctx->device);
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    return window_impl(ctx, d_frames, d_flow, d_warp, H, W, pH, pW, 0, pH * pW, nullptr, d_canvas, st);
}

int fisr_window_host(fisr_ctx* ctx, const uint8_t* h_frames, const float* h_flow, const float* h_warp, int H, int W,
                     int pH, int pW, uint8_t* h_canvas) {
    if (!ctx || !h_frames || !h_flow || !h_warp || !h_canvas) return FISR_E_INVALID;
    if (pH < 1 || pW < 1) return fail(ctx, FISR_E_INVALID, "bad tile grid");
    Guard guard(ctx->device);
    const size_t px = static_cast<size_t>(H) * W;
    const int h = H - H % (32 * pH), w = W - W % (32 * pW);
    const size_t out_bytes = static_cast<size_t>(2 * h) * (2 * w) * 9;
    int rc;
    if ((rc = ensure_stage(ctx, 1, px * 9)) != FISR_OK) return rc;
    if ((rc = ensure_stage(ctx, 2, px * 8 * 4)) != FISR_OK) return rc;
    if ((rc = ensure_stage(ctx, 3, px * 12 * 4)) != FISR_OK) return rc;
    if ((rc = ensure_stage(ctx, 4, out_bytes)) != FISR_OK) return rc;
    cudaStream_t st = ctx->stream;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->stage[1], h_frames, px * 9, cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->stage[2], h_flow, px * 8 * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->stage[3], h_warp, px * 12 * 4, cudaMemcpyHostToDevice, st));
    rc = window_impl(ctx, static_cast<const uint8_t*>(ctx->stage[1]), static_cast<const float*>(ctx->stage[2]),
                     static_cast<const float*>(ctx->stage[3]), H, W, pH, pW, 0, pH * pW,
                     static_cast<uint8_t*>(ctx->stage[4]), nullptr, st);
    if (rc != FISR_OK) return rc;
    CUDA_TRY(ctx, cudaMemcpyAsync(h_canvas, ctx->stage[4], out_bytes, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return check_kernel_error(ctx);
}

int fisr_window_host_f32(fisr_ctx* ctx, const uint8_t* h_frames, const float* h_flow, const float* h_warp, int H, int W,
                         int pH, int pW, float* h_canvas) {
    if (!ctx || !h_frames || !h_flow || !h_warp || !h_canvas) return FISR_E_INVALID;
    if (pH < 1 || pW < 1) return fail(ctx, FISR_E_INVALID, "bad tile grid");
    Guard guard(ctx->device);
    const size_t px = static_cast<size_t>(H) * W;
    const int h = H - H % (32 * pH), w = W - W % (32 * pW);
    const size_t out_bytes = static_cast<size_t>(2 * h) * (2 * w) * 9 * sizeof(float);
    int rc;
    if ((rc = ensure_stage(ctx, 1, px * 9)) != FISR_OK) return rc;
    if ((rc = ensure_stage(ctx, 2, px * 8 * 4)) != FISR_OK) return rc;
    if ((rc = ensure_stage(ctx, 3, px * 12 * 4)) != FISR_OK) return rc;
    if ((rc = ensure_stage(ctx, 4, out_bytes)) != FISR_OK) return rc;
    cudaStream_t st = ctx->stream;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->stage[1], h_frames, px * 9, cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->stage[2], h_flow, px * 8 * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->stage[3], h_warp, px * 12 * 4, cudaMemcpyHostToDevice, st));
    rc = window_impl(ctx, static_cast<const uint8_t*>(ctx->stage[1]), static_cast<const float*>(ctx->stage[2]),
                     static_cast<const float*>(ctx->stage[3]), H, W, pH, pW, 0, pH * pW, nullptr,
                     static_cast<float*>(ctx->stage[4]), st);
    if (rc != FISR_OK) return rc;
    CUDA_TRY(ctx, cudaMemcpyAsync(h_canvas, ctx->stage[4], out_bytes, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return check_kernel_error(ctx);
}

static int slot_reserve(fisr_ctx* ctx, void** buf, size_t* have, size_t need) {
    if (*have >= need) return FISR_OK;
    if (*buf) cudaFree(*buf);
    *buf = nullptr; *have = 0;
    CUDA_TRY(ctx, cudaMalloc(buf, need));
    *have = need;
    return FISR_OK;
}

int fisr_window_submit(fisr_ctx* ctx, int slot, const uint8_t* h_frames, const float* h_flow, const float* h_warp, int H,
                       int W, int pH, int pW, uint8_t* h_canvas) {
    if (!ctx || !h_frames || !h_flow || !h_warp || !h_canvas) return FISR_E_INVALID;
    if (slot < 0 || slot > 1) return fail(ctx, FISR_E_INVALID, "slot must be 0 or 1");
    if (pH < 1 || pW < 1 || H < 32 * pH || W < 32 * pW) return fail(ctx, FISR_E_INVALID, "bad tile grid %dx%d for %dx%d", pH, pW, H, W);
    Guard guard(ctx->device);
    fisr_ctx::HostSlot& sl = ctx->slots[slot];
    if (sl.busy) return fail(ctx, FISR_E_INVALID, "slot %d is still in flight: call fisr_window_wait first", slot);
    if (!ctx->h2d_stream) {
        CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking));
        CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
    }
    if (!sl.h2d_done) {
        CUDA_TRY(ctx, cudaEventCreateWithFlags(&sl.h2d_done, cudaEventDisableTiming));
        CUDA_TRY(ctx, cudaEventCreateWithFlags(&sl.compute_done, cudaEventDisableTiming));
        CUDA_TRY(ctx, cudaEventCreateWithFlags(&sl.d2h_done, cudaEventDisableTiming));
        CUDA_TRY(ctx, cudaMallocHost(reinterpret_cast<void**>(&sl.h_err), sizeof(int)));
    }
    const size_t px = static_cast<size_t>(H) * W;
    const int h = H - H % (32 * pH), w = W - W % (32 * pW);
    const size_t out_bytes = static_cast<size_t>(2 * h) * (2 * w) * 9;
    const size_t need[3] = {px * 9, px * 8 * 4, px * 12 * 4};
    const void* src[3] = {h_frames, h_flow, h_warp};
    int rc;
    for (int i = 0; i < 3; ++i)
        if ((rc = slot_reserve(ctx, &sl.in[i], &sl.in_bytes[i], need[i])) != FISR_OK) return rc;
    if ((rc = slot_reserve(ctx, &sl.out, &sl.out_bytes, out_bytes)) != FISR_OK) return rc;
    // the plan must exist before anything is enqueued (plan creation synchronises)
    {
        const TileGeom g = tile_geometry(H, W, pH, pW);
        Plan* plan = nullptr;
        for (const auto& t : g.tiles) {
            int cnt = 0;
            for (const auto& u : g.tiles) cnt += (u.yhi - u.ylo == t.yhi - t.ylo && u.xhi - u.xlo == t.xhi - t.xlo) ? 1 : 0;
            if ((rc = get_plan(ctx, std::min(cnt, kMaxTiles), t.yhi - t.ylo, t.xhi - t.xlo, &plan)) != FISR_OK) return rc;
        }
    }
    // H2D: the staging inputs of this slot are free once the previous compute that read them has finished
    if (sl.used) CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->h2d_stream, sl.compute_done, 0));
    for (int i = 0; i < 3; ++i) CUDA_TRY(ctx, cudaMemcpyAsync(sl.in[i], src[i], need[i], cudaMemcpyHostToDevice, ctx->h2d_stream));
    CUDA_TRY(ctx, cudaEventRecord(sl.h2d_done, ctx->h2d_stream));
    // compute (the context stream serialises windows: they share the plan's activation workspace)
    CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, sl.h2d_done, 0));
    rc = window_impl(ctx, static_cast<const uint8_t*>(sl.in[0]), static_cast<const float*>(sl.in[1]),
                     static_cast<const float*>(sl.in[2]), H, W, pH, pW, 0, pH * pW, static_cast<uint8_t*>(sl.out), nullptr,
                     ctx->stream);
    if (rc != FISR_OK) return rc;
    CUDA_TRY(ctx, cudaEventRecord(sl.compute_done, ctx->stream));
    // D2H
    CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->d2h_stream, sl.compute_done, 0));
    CUDA_TRY(ctx, cudaMemcpyAsync(h_canvas, sl.out, out_bytes, cudaMemcpyDeviceToHost, ctx->d2h_stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(sl.h_err, ctx->d_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->d2h_stream));
    CUDA_TRY(ctx, cudaEventRecord(sl.d2h_done, ctx->d2h_stream));
    sl.busy = true;
    sl.used = true;
    return FISR_OK;
}

int fisr_window_wait(fisr_ctx* ctx, int slot) {
    if (!ctx || slot < 0 || slot > 1) return FISR_E_INVALID;
    Guard guard(ctx->device);
    fisr_ctx::HostSlot& sl = ctx->slots[slot];
    if (!sl.busy) return fail(ctx, FISR_E_INVALID, "slot %d has nothing in flight", slot);
    CUDA_TRY(ctx, cudaEventSynchronize(sl.d2h_done));
    sl.busy = false;
    if (*sl.h_err != 0) {
        const int code = *sl.h_err;
        cudaMemsetAsync(ctx->d_err, 0, sizeof(int), ctx->stream);
        return fail(ctx, FISR_E_KERNEL, "conv kernel pipeline time-out, barrier code %d", code);
    }
    return FISR_OK;
}

int fisr_warp_device(fisr_ctx* ctx, const uint8_t* d_yuv, const float* d_flow, float flow_scale, float* d_out, int h,
                     int w, float out_scale, void* stream) {
    if (!ctx || !d_yuv || !d_flow || !d_out || h < 1 || w < 1) return FISR_E_INVALID;
    Guard guard(ctx->device);
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    launch_warp_yuv(d_yuv, d_flow, nullptr, 1, flow_scale, d_out, h, w, out_scale, st);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return FISR_OK;
}

int fisr_warp_batch_device(fisr_ctx* ctx, const uint8_t* d_frames, const float* d_flow, const int* d_src_index, int jobs,
                           float flow_scale, float* d_out, int h, int w, float out_scale, void* stream) {
    if (!ctx || !d_frames || !d_flow || !d_out || jobs < 1 || jobs > 65535 || h < 1 || w < 1) return FISR_E_INVALID;
    Guard guard(ctx->device);
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    launch_warp_yuv(d_frames, d_flow, d_src_index, jobs, flow_scale, d_out, h, w, out_scale, st);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return FISR_OK;
}

int fisr_warp_host(fisr_ctx* ctx, const uint8_t* h_yuv, const float* h_flow, float flow_scale, float* h_out, int h, int w,
                   float out_scale) {
    if (!ctx || !h_yuv || !h_flow || !h_out || h < 1 || w < 1) return FISR_E_INVALID;
    Guard guard(ctx->device);
    const size_t px = static_cast<size_t>(h) * w;
    int rc;
    if ((rc = ensure_stage(ctx, 5, px * 3)) != FISR_OK) return rc;
    if ((rc = ensure_stage(ctx, 6, px * 2 * 4)) != FISR_OK) return rc;
    if ((rc = ensure_stage(ctx, 7, px * 3 * 4)) != FISR_OK) return rc;
    cudaStream_t st = ctx->stream;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->stage[5], h_yuv, px * 3, cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->stage[6], h_flow, px * 8, cudaMemcpyHostToDevice, st));
    launch_warp_yuv(static_cast<const uint8_t*>(ctx->stage[5]), static_cast<const float*>(ctx->stage[6]), nullptr, 1, flow_scale,
                    static_cast<float*>(ctx->stage[7]), h, w, out_scale, st);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaMemcpyAsync(h_out, ctx->stage[7], px * 12, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return FISR_OK;
}

int fisr_conv3x3(fisr_ctx* ctx, const float* d_x, const float* d_w, const float* d_b, const float* d_res, int N, int H,
                 int W, int Cin, int Cout, int relu, int d2s, float* d_raw, float* d_act) {
    if (!ctx || !d_x || !d_w || !d_b || N < 1 || H < 1 || W < 1 || Cin < 1 || Cout < 1) return FISR_E_INVALID;
    if (Cout > 16 && Cout % 64) return fail(ctx, FISR_E_INVALID, "Cout must be <= 16 or a multiple of 64");
    if (d2s && Cout != 256) return fail(ctx, FISR_E_INVALID, "depth_to_space epilogue needs Cout = 256");
    if (Cout <= 16 && d_res) return fail(ctx, FISR_E_INVALID, "narrow outputs take no residual");
    if (d2s && (d_raw || d_res)) return fail(ctx, FISR_E_INVALID, "the depth_to_space epilogue has neither residual input nor fp32 output");
    Guard guard(ctx->device);
    cudaStream_t st = ctx->stream;
    Plan tmp;                       // owns the scratch buffers of this call
    tmp.planes = ctx->planes;
    Builder b{ctx, &tmp};
    ConvParam p;
    p.cin = Cin; p.cout = Cout; p.KB = (Cin + 63) / 64; p.cout_pad = Cout <= 16 ? 16 : Cout;
    p.d_w = const_cast<float*>(d_w);
    p.d_b = static_cast<float*>(b.alloc(p.cout_pad * 4, true));
    p.d_wp = static_cast<__half*>(b.alloc(static_cast<size_t>(2) * p.KB * 9 * p.cout_pad * 64 * 2, false));
    const int cs = p.KB * 64;
    ActBuf xin = b.act(N, H, W, cs);
    const int oc = d2s ? Cout / 4 : Cout, oH = d2s ? 2 * H : H, oW = d2s ? 2 * W : W;
    const int ocs = Cout <= 16 ? (ctx->planes == 3 ? 64 : 16) : oc;      // the 8-bit plane is organised in 64-channel blocks
    ActBuf yact = b.act(N, oH, oW, ocs, true);
    if (b.rc != FISR_OK) return b.rc;
    CUDA_TRY(ctx, cudaMemcpyAsync(p.d_b, d_b, Cout * 4, cudaMemcpyDeviceToDevice, st));
    launch_prep_weights(p.d_w, p.d_wp, Cin, Cout, p.KB, p.cout_pad, ctx->planes, st);
    launch_act_from_f32(d_x, Cin, xin, cs, static_cast<size_t>(N) * H * W, ctx->planes, st);
    Builder::ConvOut o;
    o.res = d_res; o.res_cs = Cout;
    o.raw = d_raw; o.raw_cs = Cout;
    o.act = (d_act || Cout > 16) ? yact : ActBuf{}; o.act_cs = ocs;
    o.relu = relu != 0; o.d2s = d2s != 0; o.scalar = Cout <= 16;
    b.conv(p, xin, cs, 0, N, H, W, o, "test");
    if (b.rc != FISR_OK) return b.rc;
    int rc = run_ops(ctx, &tmp, st);
    ctx->launches += 3;
    if (rc != FISR_OK) return rc;
    if (d_act) launch_act_to_f32(yact, ocs, 0, d_act, oc, static_cast<size_t>(N) * oH * oW, ctx->planes, st);
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return check_kernel_error(ctx);
}

int fisr_dgrad3x3(fisr_ctx* ctx, const float* d_dy, const float* d_w, const float* d_mask, const float* d_res, int N, int H,
                  int W, int Cin, int Cout, int s2d, float* d_raw, float* d_act) {
    if (!ctx || !d_dy || !d_w || N < 1 || H < 1 || W < 1 || Cin < 1 || Cout < 1) return FISR_E_INVALID;
    if (ctx->planes != 2) return fail(ctx, FISR_E_INVALID, "the backward kernels need precision f16x3");
    if (s2d && (Cin != 64 || !d_mask || d_raw || d_res)) return fail(ctx, FISR_E_INVALID, "space-to-depth dgrad: Cin = 64, gated, act output only");
    Guard guard(ctx->device);
    cudaStream_t st = ctx->stream;
    Plan tmp;
    tmp.planes = 2;
    Builder b{ctx, &tmp};
    ConvParam p;
    p.cin = Cin; p.cout = Cout; p.KB = (Cin + 63) / 64; p.cin_pad = p.KB * 64; p.OBk = (Cout + 63) / 64;
    p.d_wpT = static_cast<__half*>(b.alloc(static_cast<size_t>(2) * p.OBk * 9 * p.cin_pad * 64 * 2, false));
    const size_t npix = static_cast<size_t>(N) * H * W;
    ActBuf gin = b.act(N, H, W, p.OBk * 64), mask = b.act(N, H, W, p.cin_pad);
    const int oH = s2d ? H / 2 : H, oW = s2d ? W / 2 : W, ocs = s2d ? 4 * p.cin_pad : p.cin_pad;
    ActBuf yact = b.act(N, oH, oW, ocs, true);
    float* raw = d_raw ? b.f32(N, H, W, p.cin_pad) : nullptr;
    float* res = d_res ? b.f32(N, H, W, p.cin_pad) : nullptr;
    if (b.rc != FISR_OK) return b.rc;
    launch_prep_weights_dgrad(d_w, p.d_wpT, Cin, Cout, p.OBk, p.cin_pad, st);
    launch_act_from_f32(d_dy, Cout, gin, p.OBk * 64, npix, 2, st);
    if (d_mask) launch_act_from_f32(d_mask, Cin, mask, p.cin_pad, npix, 2, st);
    if (d_res) {
        CUDA_TRY(ctx, cudaMemsetAsync(res, 0, npix * p.cin_pad * 4, st));
        CUDA_TRY(ctx, cudaMemcpy2DAsync(res, p.cin_pad * 4, d_res, Cin * 4, Cin * 4, npix, cudaMemcpyDeviceToDevice, st));
    }
    Builder::ConvOut o;
    o.relu = false;
    o.act = yact; o.act_cs = ocs; o.s2d = s2d != 0;
    o.raw = raw; o.raw_cs = p.cin_pad;
    o.res = res; o.res_cs = p.cin_pad;
    if (d_mask) { o.mask = mask; o.mask_cs = p.cin_pad; }
    Op op{};
    if (!b.make_conv(b.bwd_view(p), gin, p.OBk * 64, 0, N, H, W, o, "test dgrad", &op)) return b.rc;
    CUDA_TRY(ctx, launch_conv3x3(op.conv, ctx->num_sms, st));
    ctx->launches += 4;
    if (d_act) launch_act_to_f32(yact, ocs, 0, d_act, s2d ? 4 * Cin : Cin, static_cast<size_t>(N) * oH * oW, 2, st);
    if (d_raw) CUDA_TRY(ctx, cudaMemcpy2DAsync(d_raw, Cin * 4, raw, p.cin_pad * 4, Cin * 4, npix, cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return check_kernel_error(ctx);
}

int fisr_wgrad3x3(fisr_ctx* ctx, const float* d_x, const float* d_dy, int N, int H, int W, int Cin, int Cout, float scale,
                  float* d_gw, float* d_gb) {
    if (!ctx || !d_x || !d_dy || !d_gw || N < 1 || H < 1 || W < 1 || Cin < 1 || Cout < 1) return FISR_E_INVALID;
    if (ctx->planes != 2) return fail(ctx, FISR_E_INVALID, "the backward kernels need precision f16x3");
    Guard guard(ctx->device);
    cudaStream_t st = ctx->stream;
    Plan tmp;
    tmp.planes = 2;
    Builder b{ctx, &tmp};
    const int CB = (Cin + 63) / 64, OB = (Cout + 63) / 64;
    ActBuf xin = b.act(N, H, W, CB * 64), dy = b.act(N, H, W, OB * 64);
    WgradLaunch L{};
    plan_wgrad(N, H, W, CB, OB, ctx->num_sms, ctx->wgrad_exact, &L);
    float* partial = static_cast<float*>(b.alloc(L.partial_floats * 4, false));
    const size_t npix = static_cast<size_t>(N) * H * W;
    if (b.rc != FISR_OK) return b.rc;
    launch_act_from_f32(d_x, Cin, xin, CB * 64, npix, 2, st);
    launch_act_from_f32(d_dy, Cout, dy, OB * 64, npix, 2, st);
    L.args.partial = partial; L.args.bias_partial = d_gb ? partial + L.bias_offset : nullptr;
    L.args.err = ctx->d_err; L.args.x_coff = 0; L.args.dy_coff = 0;
    if (!b.encode_act(&L.tmX_hi, xin.p, CB * 64, N, H, W, kWgTW + 2, kWgTH + 2)) return b.rc;
    if (!b.encode_act(&L.tmX_lo, xin.p + xin.plane, CB * 64, N, H, W, kWgTW + 2, kWgTH + 2)) return b.rc;
    if (!b.encode_act(&L.tmD_hi, dy.p, OB * 64, N, H, W, kWgTW, kWgTH)) return b.rc;
    if (!b.encode_act(&L.tmD_lo, dy.p + dy.plane, OB * 64, N, H, W, kWgTW, kWgTH)) return b.rc;
    CUDA_TRY(ctx, launch_wgrad3x3(L, st));
    launch_wgrad_reduce(L, partial, Cin, Cout, scale, d_gw, d_gb, st);
    ctx->launches += 4;
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return check_kernel_error(ctx);
}

int fisr_debug_conv_output(fisr_ctx* ctx, const char* conv_name, float* h_dst, size_t count) {
    if (!ctx || !conv_name || !h_dst) return FISR_E_INVALID;
    if (!ctx->last_plan) return fail(ctx, FISR_E_INVALID, "no forward has run yet");
    Guard guard(ctx->device);
    auto it = ctx->last_plan->debug.find(conv_name);
    if (it == ctx->last_plan->debug.end()) return fail(ctx, FISR_E_INVALID, "conv %s keeps no fp32 output", conv_name);
    const DebugTensor& t = it->second;
    const size_t n = static_cast<size_t>(t.N) * t.H * t.W * t.C;
    if (count != n) return fail(ctx, FISR_E_INVALID, "%s has %zu elements, got %zu", conv_name, n, count);
    CUDA_TRY(ctx, cudaDeviceSynchronize());
    CUDA_TRY(ctx, cudaMemcpy(h_dst, t.raw, n * 4, cudaMemcpyDeviceToHost));
    return FISR_OK;
}

int fisr_profile_ops(fisr_ctx* ctx, int N, int H, int W, int reps, int max_ops, float* ms, double* flops, double* bytes,
                     int* kinds, char* names, int name_stride) {
    if (!ctx || reps < 1) return FISR_E_INVALID;
    Guard guard(ctx->device);
    Plan* plan = nullptr;
    int rc = get_plan(ctx, N, H, W, &plan);
    if (rc != FISR_OK) return rc;
    const int n = static_cast<int>(plan->ops.size());
    if (!ms) return n;
    if (max_ops < n) return fail(ctx, FISR_E_INVALID, "plan has %d ops, buffers hold %d", n, max_ops);
    std::vector<cudaEvent_t> marks(n + 1);
    for (auto& e : marks) CUDA_TRY(ctx, cudaEventCreate(&e));
    for (int i = 0; i < n; ++i) ms[i] = 0.f;
    rc = run_ops(ctx, plan, ctx->stream);                       // warm-up
    for (int r = 0; r < reps && rc == FISR_OK; ++r) {
        rc = run_ops(ctx, plan, ctx->stream, &marks);
        ctx->launches += n;
        if (rc != FISR_OK) break;
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        for (int i = 0; i < n; ++i) {
            float t = 0.f;
            cudaEventElapsedTime(&t, marks[i], marks[i + 1]);
            ms[i] += t / reps;
        }
    }
    for (auto& e : marks) cudaEventDestroy(e);
    if (rc != FISR_OK) return rc;
    for (int i = 0; i < n; ++i) {
        const Op& op = plan->ops[i];
        if (flops) flops[i] = op.flops;
        if (bytes) bytes[i] = op.bytes;
        if (kinds) kinds[i] = static_cast<int>(op.kind) * 1000 + (op.kind == OP_CONV ? op.conv.NT * 1 + op.conv.chunks * 0 : 0);
        if (names && name_stride > 0) snprintf(names + static_cast<size_t>(i) * name_stride, name_stride, "%s", op.name);
    }
    rc = check_kernel_error(ctx);
    return rc == FISR_OK ? n : rc;
}

// ---------------------------------------------------------------- training-side entry points (forward half)
int fisr_groups2ovlp(fisr_ctx* ctx, const float* d_pred, int B, int H, int W, float* d_out, void* stream) {
    if (!ctx || !d_pred || !d_out || B < 1 || H < 1 || W < 1) return FISR_E_INVALID;
    Guard guard(ctx->device);
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    launch_groups2ovlp(d_pred, d_out, B, H, W, st);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return FISR_OK;
}

static int loss_impl(fisr_ctx* ctx, const float* const pred[3], const float* d_label, int B, int h, int w,
                     const float* lambdas, float* h_out, cudaStream_t st) {
    LossScales sc;
    const size_t ws = temporal_loss_workspace(B, h, w, &sc);
    int rc = ensure_stage(ctx, 5, ws);
    if (rc != FISR_OK) return rc;
    if (!ctx->d_scalars) CUDA_TRY(ctx, cudaMalloc(&ctx->d_scalars, 11 * sizeof(float)));
    LossLambdas lam{1.f, 1.f, 0.1f, 1.f, 0.1f, 1.f};                 // main.py:80-85
    if (lambdas) lam = LossLambdas{lambdas[0], lambdas[1], lambdas[2], lambdas[3], lambdas[4], lambdas[5]};
    launch_temporal_loss(pred, d_label, B, h, w, lam, static_cast<double*>(ctx->stage[5]), ctx->d_scalars, st);
    ctx->launches += 4;
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaMemcpyAsync(h_out, ctx->d_scalars, 11 * sizeof(float), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return FISR_OK;
}

int fisr_temporal_loss(fisr_ctx* ctx, const float* d_pred_l1, const float* d_pred_l2, const float* d_pred_l3,
                       const float* d_label, int B, int h, int w, const float* lambdas, float* h_out, void* stream) {
    if (!ctx || !d_pred_l1 || !d_pred_l2 || !d_pred_l3 || !d_label || !h_out || B < 1 || h < 4 || w < 4 || h % 4 || w % 4)
        return FISR_E_INVALID;
    Guard guard(ctx->device);
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    const float* pred[3] = {d_pred_l1, d_pred_l2, d_pred_l3};
    return loss_impl(ctx, pred, d_label, B, h, w, lambdas, h_out, st);
}

int fisr_train_forward(fisr_ctx* ctx, const float* d_data, const float* d_flow, const float* d_flow_ss2, const float* d_warp,
                       const float* d_warp_ss2, const float* d_label, int B, int h, int w, const float* lambdas, float* h_out,
                       void* stream) {
    if (!ctx || !d_data || !d_flow || !d_flow_ss2 || !d_warp || !d_warp_ss2 || !d_label || !h_out || B < 1) return FISR_E_INVALID;
    Guard guard(ctx->device);
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    Plan* plan = nullptr;
    int rc = get_plan(ctx, 4 * B, h, w, &plan);                       // the 4 weight-shared passes as one batch
    if (rc != FISR_OK) return rc;
    const size_t in_bytes = static_cast<size_t>(4) * B * h * w * IN_CH * 4;
    if ((rc = ensure_stage(ctx, 0, in_bytes)) != FISR_OK) return rc;
    launch_assemble_passes(d_data, d_flow, d_flow_ss2, d_warp, d_warp_ss2, static_cast<float*>(ctx->stage[0]), B, h, w, st);
    launch_pack_input(static_cast<const float*>(ctx->stage[0]), 4 * B, h, w, IN_CH, plan->in_lvl[2], plan->in_lvl[1],
                      plan->in_lvl[0], plan->planes, st);
    ctx->launches += 2;
    rc = run_plan(ctx, plan, st);
    if (rc != FISR_OK) return rc;
    const float* pred[3] = {plan->pred[0], plan->pred[1], plan->pred[2]};
    rc = loss_impl(ctx, pred, d_label, B, h, w, lambdas, h_out, st);
    if (rc != FISR_OK) return rc;
    return check_kernel_error(ctx);
}

int fisr_adam_step(fisr_ctx* ctx, const float* const* d_grads, int n_grads, float lr, float beta1, float beta2, float eps) {
    if (!ctx || !d_grads) return FISR_E_INVALID;
    if (n_grads != 2 * static_cast<int>(ctx->params.size()))
        return fail(ctx, FISR_E_INVALID, "expected %d gradient tensors (creation order, w then b), got %d", 2 * (int)ctx->params.size(), n_grads);
    Guard guard(ctx->device);
    std::vector<const float*> g(d_grads, d_grads + n_grads);
    for (int i = 0; i < n_grads; ++i)
        if (!g[i]) return fail(ctx, FISR_E_INVALID, "gradient %d is NULL", i);
    CUDA_TRY(ctx, cudaDeviceSynchronize());          // the caller's gradients may come from any stream
    const int rc = adam_impl(ctx, g, lr, beta1, beta2, eps, ctx->stream);
    if (rc != FISR_OK) return rc;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return FISR_OK;
}

// Adam slot variables "<var>/Adam" (m) and "<var>/Adam_1" (v) of the reference's training checkpoints (tf.train.Saver
// saves them next to the weights, FISRnet.py:1092-1099): which = 0 -> m, 1 -> v.
static int adam_slot_ptr(fisr_ctx* ctx, const char* name, int which, size_t count, float** out) {
    int ci; bool is_w;
    int rc = split_param_name(ctx, name, &ci, &is_w);
    if (rc != FISR_OK) return rc;
    if (which != 0 && which != 1) return fail(ctx, FISR_E_INVALID, "Adam slot must be 0 (m) or 1 (v)");
    ConvParam& p = ctx->params[ci];
    const size_t expect = is_w ? static_cast<size_t>(9) * p.cin * p.cout : static_cast<size_t>(p.cout);
    if (count != expect) return fail(ctx, FISR_E_INVALID, "%s has %zu elements, got %zu", name, expect, count);
    if ((rc = ensure_adam_slots(ctx, ctx->stream)) != FISR_OK) return rc;
    *out = is_w ? (which ? p.v_w : p.m_w) : (which ? p.v_b : p.m_b);
    return FISR_OK;
}

int fisr_get_adam_slot(fisr_ctx* ctx, const char* name, int which, float* h_data, size_t count) {
    if (!ctx || !h_data) return FISR_E_INVALID;
    Guard guard(ctx->device);
    float* src = nullptr;
    const int rc = adam_slot_ptr(ctx, name, which, count, &src);
    if (rc != FISR_OK) return rc;
    CUDA_TRY(ctx, cudaDeviceSynchronize());
    CUDA_TRY(ctx, cudaMemcpy(h_data, src, count * 4, cudaMemcpyDeviceToHost));
    return FISR_OK;
}

int fisr_set_adam_slot(fisr_ctx* ctx, const char* name, int which, const float* h_data, size_t count) {
    if (!ctx || !h_data) return FISR_E_INVALID;
    Guard guard(ctx->device);
    float* dst = nullptr;
    const int rc = adam_slot_ptr(ctx, name, which, count, &dst);
    if (rc != FISR_OK) return rc;
    CUDA_TRY(ctx, cudaDeviceSynchronize());
    CUDA_TRY(ctx, cudaMemcpy(dst, h_data, count * 4, cudaMemcpyHostToDevice));
    return FISR_OK;
}

int fisr_adam_set_steps(fisr_ctx* ctx, long long step) {
    if (!ctx || step < 0) return FISR_E_INVALID;
    ctx->adam_t = step;
    return FISR_OK;
}

long long fisr_adam_steps(const fisr_ctx* ctx) { return ctx ? ctx->adam_t : 0; }

int fisr_adam_reset(fisr_ctx* ctx, long long step) {
    if (!ctx || step < 0) return FISR_E_INVALID;
    Guard guard(ctx->device);
    CUDA_TRY(ctx, cudaDeviceSynchronize());
    for (auto& p : ctx->params) {
        cudaFree(p.m_w); cudaFree(p.v_w); cudaFree(p.m_b); cudaFree(p.v_b);
        p.m_w = p.v_w = p.m_b = p.v_b = nullptr;
    }
    ctx->adam_t = step;
    return FISR_OK;
}

// ---------------------------------------------------------------- training step (forward + loss + backward [+ Adam])
static int train_backward_impl(fisr_ctx* ctx, const float* d_data, const float* d_flow, const float* d_flow_ss2,
                               const float* d_warp, const float* d_warp_ss2, const float* d_label, int B, int h, int w,
                               const float* lambdas, float* h_out, cudaStream_t st) {
    Plan* plan = nullptr;
    int rc = get_plan(ctx, 4 * B, h, w, &plan);
    if (rc != FISR_OK) return rc;
    if ((rc = ensure_train_params(ctx, ctx->stream)) != FISR_OK) return rc;
    if ((rc = ensure_backward(ctx, plan, B)) != FISR_OK) return rc;
    if (st != ctx->stream) CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));     // operand packing above
    const size_t in_bytes = static_cast<size_t>(4) * B * h * w * IN_CH * 4;
    if ((rc = ensure_stage(ctx, 0, in_bytes)) != FISR_OK) return rc;
    launch_assemble_passes(d_data, d_flow, d_flow_ss2, d_warp, d_warp_ss2, static_cast<float*>(ctx->stage[0]), B, h, w, st);
    launch_pack_input(static_cast<const float*>(ctx->stage[0]), 4 * B, h, w, IN_CH, plan->in_lvl[2], plan->in_lvl[1],
                      plan->in_lvl[0], plan->planes, st);
    ctx->launches += 2;
    if ((rc = run_plan(ctx, plan, st)) != FISR_OK) return rc;
    plan->label = d_label;
    plan->lam = LossLambdas{1.f, 1.f, 0.1f, 1.f, 0.1f, 1.f};
    if (lambdas) plan->lam = LossLambdas{lambdas[0], lambdas[1], lambdas[2], lambdas[3], lambdas[4], lambdas[5]};
    if ((rc = run_backward(ctx, plan, st)) != FISR_OK) return rc;
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_gmax, 0, sizeof(unsigned), st));
    {   // max |g| over every gradient tensor in one launch (overflow probe of the loss scale)
        std::vector<const float*> g;
        for (auto& p : ctx->params) { g.push_back(p.g_w); g.push_back(p.g_b); }
        unsigned blocks = 0;
        if ((rc = upload_mt_table(ctx, g, st, &blocks)) != FISR_OK) return rc;
        launch_mt_absmax(ctx->d_mt, static_cast<int>(ctx->mt_host.size()), blocks, ctx->d_gmax, st);
        ctx->launches++;
    }
    const float* pred[3] = {plan->pred[0], plan->pred[1], plan->pred[2]};
    float scalars[11];
    if ((rc = loss_impl(ctx, pred, d_label, B, h, w, lambdas, scalars, st)) != FISR_OK) return rc;    // synchronises
    if (h_out) memcpy(h_out, scalars, sizeof scalars);
    unsigned gmax = 0;
    CUDA_TRY(ctx, cudaMemcpy(&gmax, ctx->d_gmax, sizeof gmax, cudaMemcpyDeviceToHost));
    if ((rc = check_kernel_error(ctx)) != FISR_OK) return rc;
    if (gmax >= 0x7F800000u)
        return fail(ctx, FISR_E_OVERFLOW, "non-finite gradient (loss scale %g overflowed the fp16 gradient planes): lower it with fisr_set_loss_scale", plan->loss_scale);
    return FISR_OK;
}

int fisr_train_backward(fisr_ctx* ctx, const float* d_data, const float* d_flow, const float* d_flow_ss2, const float* d_warp,
                        const float* d_warp_ss2, const float* d_label, int B, int h, int w, const float* lambdas, float* h_out,
                        void* stream) {
    if (!ctx || !d_data || !d_flow || !d_flow_ss2 || !d_warp || !d_warp_ss2 || !d_label || B < 1) return FISR_E_INVALID;
    Guard guard(ctx->device);
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    return train_backward_impl(ctx, d_data, d_flow, d_flow_ss2, d_warp, d_warp_ss2, d_label, B, h, w, lambdas, h_out, st);
}

static int adam_apply_impl(fisr_ctx* ctx, float lr, float beta1, float beta2, float eps, cudaStream_t st) {
    std::vector<const float*> g;
    for (auto& p : ctx->params) {
        if (!p.g_w) return fail(ctx, FISR_E_INVALID, "no gradients yet: call fisr_train_backward first");
        g.push_back(p.g_w);
        g.push_back(p.g_b);
    }
    return adam_impl(ctx, g, lr, beta1, beta2, eps, st);
}

int fisr_adam_apply(fisr_ctx* ctx, float lr, float beta1, float beta2, float eps) {
    if (!ctx) return FISR_E_INVALID;
    Guard guard(ctx->device);
    // fisr_train_backward is synchronous, so the gradients are complete; later calls may use another stream: finish here
    const int rc = adam_apply_impl(ctx, lr, beta1, beta2, eps, ctx->stream);
    if (rc != FISR_OK) return rc;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return FISR_OK;
}

// One `sess.run(optim)`.  Dynamic loss scaling: when the fp16 gradient planes overflow, the update is skipped, the scale is
// lowered (to the default, then by 16 per retry, up to 5 retries) and the step is retried; the lowered scale stays in force for the following steps.
int fisr_train_step(fisr_ctx* ctx, const float* d_data, const float* d_flow, const float* d_flow_ss2, const float* d_warp,
                    const float* d_warp_ss2, const float* d_label, int B, int h, int w, const float* lambdas, float lr,
                    float* h_out, void* stream) {
    if (!ctx) return FISR_E_INVALID;
    int rc = FISR_OK;
    for (int attempt = 0; attempt < 6; ++attempt) {
        rc = fisr_train_backward(ctx, d_data, d_flow, d_flow_ss2, d_warp, d_warp_ss2, d_label, B, h, w, lambdas, h_out, stream);
        if (rc != FISR_E_OVERFLOW || attempt == 5) break;
        const float cur = fisr_get_loss_scale(ctx, B, h, w);
        const float def = static_cast<float>(std::exp2(std::floor(std::log2(static_cast<double>(B) * (2.0 * h) * (2.0 * w) * 3.0))));
        // an override above the default falls back to the default first, then the scale drops by 16 per retry
        if ((rc = fisr_set_loss_scale(ctx, cur > def ? def : cur / 16.f)) != FISR_OK) return rc;
    }
    if (rc != FISR_OK) return rc;
    Guard guard(ctx->device);
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    return adam_apply_impl(ctx, lr, 0.9f, 0.999f, 1e-8f, st);      // asynchronous on the step's stream: 2 launches, no host sync
}

int fisr_get_grad(fisr_ctx* ctx, const char* name, float* h_data, size_t count) {
    if (!ctx || !h_data) return FISR_E_INVALID;
    int ci; bool is_w;
    int rc = split_param_name(ctx, name, &ci, &is_w);
    if (rc != FISR_OK) return rc;
    Guard guard(ctx->device);
    ConvParam& p = ctx->params[ci];
    if (!p.g_w) return fail(ctx, FISR_E_INVALID, "no gradients yet: call fisr_train_backward first");
    const size_t expect = is_w ? static_cast<size_t>(9) * p.cin * p.cout : static_cast<size_t>(p.cout);
    if (count != expect) return fail(ctx, FISR_E_INVALID, "%s has %zu elements, got %zu", name, expect, count);
    CUDA_TRY(ctx, cudaDeviceSynchronize());
    CUDA_TRY(ctx, cudaMemcpy(h_data, is_w ? p.g_w : p.g_b, count * 4, cudaMemcpyDeviceToHost));
    return FISR_OK;
}

int fisr_set_loss_scale(fisr_ctx* ctx, float scale) {
    if (!ctx || !(scale >= 0.f)) return FISR_E_INVALID;
    ctx->loss_scale_override = scale;
    for (auto& kv : ctx->plans)
        if (kv.second->has_bwd && scale > 0.f) kv.second->loss_scale = scale;
    return FISR_OK;
}

int fisr_set_wgrad_exact(fisr_ctx* ctx, int exact) {
    if (!ctx) return FISR_E_INVALID;
    if ((exact != 0) != ctx->wgrad_exact) {
        Guard guard(ctx->device);
        cudaDeviceSynchronize();
        ctx->plans.clear();               // backward launch lists bake the choice in
        ctx->last_plan = nullptr;
        ctx->wgrad_exact = exact != 0;
    }
    return FISR_OK;
}

float fisr_get_loss_scale(fisr_ctx* ctx, int B, int h, int w) {
    if (!ctx) return 0.f;
    if (ctx->loss_scale_override > 0.f) return ctx->loss_scale_override;
    return std::exp2(std::floor(std::log2(static_cast<double>(B) * (2.0 * h) * (2.0 * w) * 3.0)));
}

int fisr_profile_train(fisr_ctx* ctx, int B, int h, int w, int reps, int max_ops, float* ms, double* flops, char* names,
                       int name_stride) {
    if (!ctx || reps < 1 || B < 1) return FISR_E_INVALID;
    Guard guard(ctx->device);
    Plan* plan = nullptr;
    int rc = get_plan(ctx, 4 * B, h, w, &plan);
    if (rc != FISR_OK) return rc;
    if ((rc = ensure_train_params(ctx, ctx->stream)) != FISR_OK) return rc;
    if ((rc = ensure_backward(ctx, plan, B)) != FISR_OK) return rc;
    const int n = static_cast<int>(plan->bwd.size());
    if (!ms) return n;
    if (max_ops < n) return fail(ctx, FISR_E_INVALID, "backward has %d ops, buffers hold %d", n, max_ops);
    if (!plan->label) return fail(ctx, FISR_E_INVALID, "run fisr_train_backward once before profiling (it binds the label)");
    std::vector<cudaEvent_t> marks(n + 1);
    for (auto& e : marks) CUDA_TRY(ctx, cudaEventCreate(&e));
    for (int i = 0; i < n; ++i) ms[i] = 0.f;
    cudaStream_t st = ctx->stream;
    for (int r = 0; r < reps + 1 && rc == FISR_OK; ++r) {          // first pass = warm-up
        for (int i = 0; i < n && rc == FISR_OK; ++i) {
            cudaEventRecord(marks[i], st);
            rc = plan->bwd[i].run(st);
            ctx->launches += plan->bwd[i].launches;
        }
        cudaEventRecord(marks[n], st);
        if (rc != FISR_OK) break;
        CUDA_TRY(ctx, cudaStreamSynchronize(st));
        if (r == 0) continue;
        for (int i = 0; i < n; ++i) {
            float t = 0.f;
            cudaEventElapsedTime(&t, marks[i], marks[i + 1]);
            ms[i] += t / reps;
        }
    }
    for (auto& e : marks) cudaEventDestroy(e);
    if (rc != FISR_OK) return rc;
    for (int i = 0; i < n; ++i) {
        if (flops) flops[i] = plan->bwd[i].flops;
        if (names && name_stride > 0) snprintf(names + static_cast<size_t>(i) * name_stride, name_stride, "%s", plan->bwd[i].name.c_str());
    }
    rc = check_kernel_error(ctx);
    return rc == FISR_OK ? n : rc;
}

// ---------------------------------------------------------------- multi-GPU frame exchange over peer memory
int fisr_ipc_alloc(fisr_ctx* ctx, size_t bytes, void** d_ptr, unsigned char* handle64) {
    if (!ctx || !d_ptr || !handle64 || bytes == 0) return FISR_E_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    Guard guard(ctx->device);
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return fail(ctx, FISR_E_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    cudaIpcMemHandle_t h;
    if ((e = cudaIpcGetMemHandle(&h, p)) != cudaSuccess) {
        cudaFree(p);
        return fail(ctx, FISR_E_CUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    }
    CUDA_TRY(ctx, cudaMemset(p, 0, bytes));
    memcpy(handle64, &h, 64);
    *d_ptr = p;
    return FISR_OK;
}

int fisr_ipc_open(fisr_ctx* ctx, const unsigned char* handle64, void** d_ptr) {
    if (!ctx || !handle64 || !d_ptr) return FISR_E_INVALID;
    Guard guard(ctx->device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return fail(ctx, FISR_E_CUDA, "cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e));
    *d_ptr = p;
    return FISR_OK;
}

int fisr_ipc_close(fisr_ctx* ctx, void* d_ptr) {
    if (!ctx || !d_ptr) return FISR_E_INVALID;
    Guard guard(ctx->device);
    CUDA_TRY(ctx, cudaIpcCloseMemHandle(d_ptr));
    return FISR_OK;
}

int fisr_ipc_free(fisr_ctx* ctx, void* d_ptr) {
    if (!ctx || !d_ptr) return FISR_E_INVALID;
    Guard guard(ctx->device);
    CUDA_TRY(ctx, cudaFree(d_ptr));
    return FISR_OK;
}

int fisr_copy2d_async(fisr_ctx* ctx, void* d_dst, size_t dst_pitch, const void* d_src, size_t src_pitch, size_t width_bytes,
                      size_t rows, void* stream) {
    if (!ctx || !d_dst || !d_src || width_bytes == 0 || rows == 0 || dst_pitch < width_bytes || src_pitch < width_bytes) return FISR_E_INVALID;
    Guard guard(ctx->device);
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    CUDA_TRY(ctx, cudaMemcpy2DAsync(d_dst, dst_pitch, d_src, src_pitch, width_bytes, rows, cudaMemcpyDeviceToDevice, st));
    return FISR_OK;
}

long long fisr_launch_count(const fisr_ctx* ctx) { return ctx ? ctx->launches : 0; }

int fisr_plan_info(fisr_ctx* ctx, int N, int H, int W, double* flops, double* mma_efficiency, int* num_launches,
                   size_t* workspace_bytes) {
    if (!ctx) return FISR_E_INVALID;
    Guard guard(ctx->device);
    Plan* plan = nullptr;
    int rc = get_plan(ctx, N, H, W, &plan);
    if (rc != FISR_OK) return rc;
    if (flops) *flops = plan->flops;
    if (mma_efficiency) *mma_efficiency = plan->flops > 0 ? plan->eff_weighted / plan->flops : 0;
    if (num_launches) *num_launches = static_cast<int>(plan->ops.size());
    if (workspace_bytes) *workspace_bytes = plan->bytes;
    return FISR_OK;
}

}  // extern "C"
