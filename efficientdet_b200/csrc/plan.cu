// Plan-level C ABI: the whole EfficientDet inference path -- what the reference's callers get from
//   model, prediction_model = efficientdet(phi, ...)              model.py:356-452
//   prediction_model.predict_on_batch([images (, anchors)])       inference.py:57-59, predict.py:101-105
// -- behind six C entry points, so that a host in any language can run forward / detect without Python:
//   effdet_plan_create -> effdet_plan_weight_info / effdet_plan_bind_weights(_host) -> effdet_forward /
//   effdet_detect / effdet_detect_host -> effdet_plan_destroy.
// The plan is a C++ lowering of the network for a fixed (phi, image size, batch, classes, BiFPN kind, dtype): the
// EfficientNet block table (efficientnet.py:99-114, :191-207, :428-469), the BiFPN wiring (model.py:93-268), the
// shared heads writing straight into the concatenated outputs (model.py:271-353, :393-398) become a flat list of
// launches of THIS library's per-layer entry points over plan-owned NHWC buffers (liveness-based reuse), captured
// once into a CUDA graph.  It mirrors efficientdet_b200/engine.py line by line (same launches, same arguments):
// tests/test_gpu_plan_cabi.py checks that both produce bit-identical outputs.
// Host code only: no kernels live in this file.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <functional>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"

using namespace effdet;

namespace {

const int kWBifpn[7] = {64, 88, 112, 160, 224, 288, 384};                          // model.py:28
const double kCoeff[7][2] = {{1.0, 1.0}, {1.0, 1.1}, {1.1, 1.2}, {1.2, 1.4}, {1.4, 1.8}, {1.6, 2.2}, {1.8, 2.6}};
// (kernel, repeats, in, out, expand, stride)   efficientnet.py:99-114 (se_ratio .25, id_skip everywhere)
const int kBlocks[7][6] = {{3, 1, 32, 16, 1, 1}, {3, 2, 16, 24, 6, 2}, {5, 2, 24, 40, 6, 2}, {3, 3, 40, 80, 6, 2},
                           {5, 3, 80, 112, 6, 1}, {5, 4, 112, 192, 6, 2}, {3, 1, 192, 320, 6, 1}};
const float kBnEpsBackbone = 1e-3f, kBnEpsBifpn = 1e-4f;                            // efficientnet.py:233-236, model.py:42-45

int round_filters(int filters, double width) {                                       // efficientnet.py:191-201
    const double f = filters * width;
    int nf = (int)(f + 4) / 8 * 8;
    if (nf < 8) nf = 8;
    if (nf < 0.9 * f) nf += 8;
    return nf;
}
int round_repeats(int r, double depth) { return (int)ceil(depth * r); }             // efficientnet.py:204-207

struct Block {
    std::string prefix;
    int k, stride, cin, cmid, cout, expand, se;
    bool skip, tap;
};

struct WeightInfo {
    std::string name;
    int ndim, dims[4];
    size_t offset, count;       // into the flat fp32 buffer (elements)
};

struct Val {
    size_t bytes = 0;
    bool keep = false;
    int last_use = -1;
    void *ptr = nullptr;
};

struct Op {
    std::vector<int> ins, outs;
    std::function<int(cudaStream_t)> run;
};

}  // namespace

struct effdet_plan {
    int phi, S, B, C, weighted, dtype;
    unsigned flags;
    int W, D, hd, c0;
    std::vector<Block> blocks;
    std::vector<WeightInfo> winfo;
    std::map<std::string, int> windex;
    float *flat = nullptr;                 // all weights, fp32, Keras shapes, creation order
    size_t flat_count = 0;
    std::vector<char> bound;
    bool finalized = false;
    // derived tensors
    struct Folded { float *scale, *shift; int C; float eps; std::string bn; };
    std::vector<Folded> folded;
    std::map<std::string, int> folded_index;
    struct Panel { std::string key; int taps, cin, cout; void *ptr; bool split; };
    std::vector<Panel> panels;
    float *lut = nullptr;                  // (3,256) normalisation table of the uint8 stem
    // launch list
    std::vector<Val> vals;
    std::vector<Op> ops;
    std::vector<void *> owned;             // every cudaMalloc of the plan
    std::vector<void *> owned_host;        // dry run only (no CUDA driver + EFFDET_DRY_RUN): host stand-ins
    bool host_mem = false;
    std::vector<effdet_conv_desc *> descs;
    int v_images = -1, v_reg = -1, v_cls = -1;
    size_t N = 0;
    cudaGraphExec_t graph = nullptr;
    cudaStream_t capture_stream = nullptr;
    // detection tail
    float *anchors = nullptr, *boxes = nullptr, *user_anchors = nullptr;
    void *ws = nullptr;
    size_t ws_bytes = 0, cand_capacity = 0;
    int32_t *status = nullptr;
    float *out_boxes = nullptr, *out_scores = nullptr;
    int32_t *out_labels = nullptr;
    int out_max_det = 0;
    void *host_in = nullptr;               // pinned staging of effdet_detect_host
    void *host_out = nullptr;
    size_t host_out_bytes = 0;

    float *w(const std::string &name) const { return flat + winfo[windex.at(name)].offset; }
    size_t es() const { return dtype == EFFDET_BF16 ? 2 : 4; }
};

namespace {

int dev_alloc(effdet_plan *p, void **out, size_t bytes) {
    void *q = nullptr;
    const cudaError_t alloc_error = cudaMalloc(&q, bytes < 16 ? 16 : bytes);
    if (alloc_error != cudaSuccess && getenv("EFFDET_DRY_RUN") != nullptr && (p->host_mem || p->owned.empty())) {
        // Dry run (tests/test_dry_run_launches.py, machines without a CUDA driver): the lowering is built over host
        // memory so that effdet_plan_dry_run can walk the launch list through the entry points' host side.  Nothing
        // can execute: every launch fails at its first CUDA call.
        (void)cudaGetLastError();
        if (posix_memalign(&q, 256, (bytes + 255) / 256 * 256 + 256) != 0)
            return fail(EFFDET_E_CUDA, "effdet_plan: dry-run host allocation of %s%lld bytes failed", "", (long long)bytes);
        p->owned_host.push_back(q);
        p->host_mem = true;
        *out = q;
        return EFFDET_OK;
    }
    if (alloc_error != cudaSuccess) {
        snprintf(::effdet::g_err, sizeof(::effdet::g_err), "effdet_plan: cudaMalloc(%lld bytes) -> %s", (long long)bytes,
                 cudaGetErrorString(alloc_error));
        return EFFDET_E_CUDA;
    }
    p->owned.push_back(q);
    *out = q;
    return EFFDET_OK;
}

void add_weight(effdet_plan *p, const std::string &name, int ndim, int d0, int d1 = 1, int d2 = 1, int d3 = 1) {
    WeightInfo w;
    w.name = name; w.ndim = ndim;
    w.dims[0] = d0; w.dims[1] = d1; w.dims[2] = d2; w.dims[3] = d3;
    w.count = (size_t)d0 * d1 * d2 * d3;
    w.offset = p->flat_count;
    p->flat_count += (w.count + 3) / 4 * 4;             // 16-byte aligned slots
    p->windex[name] = (int)p->winfo.size();
    p->winfo.push_back(w);
}
void add_bn(effdet_plan *p, const std::string &name, int C, float eps) {
    add_weight(p, name + "/gamma", 1, C); add_weight(p, name + "/beta", 1, C);
    add_weight(p, name + "/moving_mean", 1, C); add_weight(p, name + "/moving_variance", 1, C);
    effdet_plan::Folded f; f.scale = f.shift = nullptr; f.C = C; f.eps = eps; f.bn = name;
    p->folded_index[name] = (int)p->folded.size();
    p->folded.push_back(f);
}
std::string fuse_name(int k) { return k == 0 ? "w_bi_fpn_add" : "w_bi_fpn_add_" + std::to_string(k); }
const char *kNodes[8] = {"U_P6", "U_P5", "U_P4", "U_P3", "D_P4", "D_P5", "D_P6", "D_P7"};

// weight manifest in the creation order of the reference graph (== efficientdet_b200/model.py _init_weights)
void build_manifest(effdet_plan *p) {
    const double wc = kCoeff[p->phi][0], dc = kCoeff[p->phi][1];
    p->c0 = round_filters(32, wc);
    add_weight(p, "stem_conv/kernel", 4, 3, 3, 3, p->c0);
    add_bn(p, "stem_bn", p->c0, kBnEpsBackbone);
    for (int idx = 0; idx < 7; ++idx) {
        const int k = kBlocks[idx][0], rep = round_repeats(kBlocks[idx][1], dc);
        const int cin = round_filters(kBlocks[idx][2], wc), cout = round_filters(kBlocks[idx][3], wc);
        const int e = kBlocks[idx][4], s = kBlocks[idx][5];
        for (int r = 0; r < rep; ++r) {
            Block b;
            b.prefix = "block" + std::to_string(idx + 1) + std::string(1, (char)('a' + r)) + "_";
            b.k = k; b.stride = r == 0 ? s : 1; b.cin = r == 0 ? cin : cout; b.cout = cout; b.expand = e;
            b.cmid = b.cin * e; b.se = b.cin / 4 > 0 ? b.cin / 4 : 1;          // max(1, int(cin * 0.25))
            b.skip = b.stride == 1 && b.cin == b.cout;
            b.tap = false;
            p->blocks.push_back(b);
        }
        if ((idx < 6 && kBlocks[idx + 1][5] == 2) || idx == 6) p->blocks.back().tap = true;   // efficientnet.py:466-469
    }
    for (const Block &b : p->blocks) {
        const std::string &q = b.prefix;
        if (b.expand != 1) { add_weight(p, q + "expand_conv/kernel", 4, 1, 1, b.cin, b.cmid); add_bn(p, q + "expand_bn", b.cmid, kBnEpsBackbone); }
        add_weight(p, q + "dwconv/depthwise_kernel", 4, b.k, b.k, b.cmid, 1);
        add_bn(p, q + "bn", b.cmid, kBnEpsBackbone);
        add_weight(p, q + "se_reduce/kernel", 4, 1, 1, b.cmid, b.se); add_weight(p, q + "se_reduce/bias", 1, b.se);
        add_weight(p, q + "se_expand/kernel", 4, 1, 1, b.se, b.cmid); add_weight(p, q + "se_expand/bias", 1, b.cmid);
        add_weight(p, q + "project_conv/kernel", 4, 1, 1, b.cmid, b.cout);
        add_bn(p, q + "project_bn", b.cout, kBnEpsBackbone);
    }
    std::vector<int> feat;
    for (const Block &b : p->blocks) if (b.tap) feat.push_back(b.cout);
    const int W = p->W;
    for (int i = 0; i < p->D; ++i) {
        const std::string pre = "BiFPN_" + std::to_string(i) + "_";
        for (int l = 3; l <= 7; ++l) {
            int cin = W, k = 1;
            if (i == 0) { cin = l <= 5 ? feat[l - 1] : (l == 6 ? feat[4] : W); k = l <= 5 ? 1 : 3; }
            const std::string n = pre + "P" + std::to_string(l);
            add_weight(p, n + "_conv/kernel", 4, k, k, cin, W);
            add_bn(p, n + "_bn", W, kBnEpsBifpn);
        }
        for (int j = 0; j < 8; ++j) {
            const std::string n = pre + kNodes[j];
            add_weight(p, n + "_dconv/depthwise_kernel", 4, 3, 3, W, 1);
            add_bn(p, n + "_bn", W, kBnEpsBifpn);
            if (p->weighted) {
                const std::string f = fuse_name(8 * i + j);
                add_weight(p, f + "/" + f, 1, (j >= 4 && j <= 6) ? 3 : 2);
            }
        }
    }
    const int A = 9;
    for (int i = 0; i < p->hd; ++i) {
        add_weight(p, "box_head/regress_head_conv_" + std::to_string(i) + "/kernel", 4, 3, 3, W, W);
        add_weight(p, "box_head/regress_head_conv_" + std::to_string(i) + "/bias", 1, W);
    }
    add_weight(p, "box_head/regress_head_conv_final/kernel", 4, 3, 3, W, A * 4);
    add_weight(p, "box_head/regress_head_conv_final/bias", 1, A * 4);
    for (int i = 0; i < p->hd; ++i) {
        add_weight(p, "class_head/class_head_" + std::to_string(i) + "/kernel", 4, 3, 3, W, W);
        add_weight(p, "class_head/class_head_" + std::to_string(i) + "/bias", 1, W);
    }
    add_weight(p, "class_head/pyramid_classification/kernel", 4, 3, 3, W, A * p->C);
    add_weight(p, "class_head/pyramid_classification/bias", 1, A * p->C);
}

// ---------------------------------------------------------------- launch-list construction
int new_val(effdet_plan *p, size_t bytes, bool keep = false) {
    Val v; v.bytes = bytes; v.keep = keep;
    p->vals.push_back(v);
    return (int)p->vals.size() - 1;
}
int act_val(effdet_plan *p, int H, int Wd, int C, bool keep = false) {
    return new_val(p, (size_t)p->B * H * Wd * C * p->es(), keep);
}
void add_op(effdet_plan *p, std::vector<int> ins, std::vector<int> outs, std::function<int(cudaStream_t)> run) {
    Op op; op.ins = std::move(ins); op.outs = std::move(outs); op.run = std::move(run);
    p->ops.push_back(std::move(op));
}
void *static_panel(effdet_plan *p, const std::string &key, int taps, int cin, int cout, bool split = false) {
    for (auto &q : p->panels) if (q.key == key) return q.ptr;
    effdet_plan::Panel q; q.key = key; q.taps = taps; q.cin = cin; q.cout = cout; q.ptr = nullptr; q.split = split;
    const size_t elems = split ? effdet_conv_weight_panel_split_elems(taps, cin, cout)
                               : effdet_conv_weight_panel_elems(taps, cin, cout);
    if (dev_alloc(p, &q.ptr, elems * 2) != EFFDET_OK) return nullptr;
    p->panels.push_back(q);
    return q.ptr;
}

struct ConvArgs {
    std::vector<int> xs, ys, residuals;     // Val indices (residuals: -1 = none)
    std::vector<int> H, Wd;
    std::vector<int> ldc;
    std::vector<long long> ybs;
    std::vector<size_t> y_off;              // byte offsets into ys[i]
    std::string weight;
    int cin, cout, k = 1, stride = 1;
    const float *scale = nullptr, *shift = nullptr;
    int act = EFFDET_ACT_NONE, gate = -1;
    int in_dtype = -1, out_dtype = -1;
    bool pre_split = false;                 // xs already are the (B,H,W,2*cin) bf16 hi | lo planes of an fp32 tensor
};

// engine.Plan.conv: effdet_conv2d over 1..5 groups sharing the weights; tcgen05 path for bf16 activations
int emit_conv(effdet_plan *p, const ConvArgs &a) {
    const int n = (int)a.xs.size();
    const int in_dt = a.in_dtype < 0 ? p->dtype : a.in_dtype, out_dt = a.out_dtype < 0 ? p->dtype : a.out_dtype;
    bool use_tc = in_dt == EFFDET_BF16 && a.cin % 8 == 0 && (a.stride == 1 || (a.stride == 2 && a.gate < 0));
    // fp32 accuracy mode: the tensor-core kernel on the bf16 hi | lo split of the fp32 activations (engine.py
    // Plan.conv `split`; EFFDET_FP32_TC=0 keeps the exact SIMT kernel)
    static const bool fp32_tc = !(getenv("EFFDET_FP32_TC") && atoi(getenv("EFFDET_FP32_TC")) == 0);
    const bool split = a.pre_split || (fp32_tc && in_dt == EFFDET_F32 && out_dt == EFFDET_F32 && a.cin % 8 == 0 &&
                                       (a.stride == 1 || (a.stride == 2 && a.gate < 0)));
    int gate_panel = -1;
    void *panel = nullptr;
    std::vector<int> xs_used = a.xs;
    if (split) {
        for (int i = 0; i < n && !a.pre_split; ++i) {
            const int xf = a.xs[i], H = a.H[i], Wd = a.Wd[i], cin = a.cin, B = p->B;
            const int xv = new_val(p, (size_t)B * H * Wd * 2 * cin * 2);
            xs_used[i] = xv;
            add_op(p, {xf}, {xv}, [=](cudaStream_t st) {
                return effdet_split_bf16((const float *)p->vals[xf].ptr, p->vals[xv].ptr, (size_t)B * H * Wd, cin, st);
            });
        }
        if (a.gate >= 0) {
            gate_panel = new_val(p, effdet_conv_weight_panel_split_elems(p->B, a.cin, a.cout) * 2);
            const int gv = a.gate, gp = gate_panel, cin = a.cin, cout = a.cout, B = p->B;
            const std::string wk = a.weight;
            add_op(p, {gv}, {gp}, [=](cudaStream_t st) {
                return effdet_conv_weight_panel_split(p->w(wk), p->vals[gp].ptr, 1, cin, cout, (const float *)p->vals[gv].ptr, B, st);
            });
        } else {
            panel = static_panel(p, a.weight, a.k * a.k, a.cin, a.cout, true);
            if (!panel) return EFFDET_E_CUDA;
        }
        use_tc = true;
    } else if (use_tc && a.gate >= 0) {
        gate_panel = new_val(p, effdet_conv_weight_panel_elems(p->B, a.cin, a.cout) * 2);
        const int gv = a.gate, gp = gate_panel, cin = a.cin, cout = a.cout, B = p->B;
        const std::string wk = a.weight;
        add_op(p, {gv}, {gp}, [=](cudaStream_t st) {
            return effdet_conv_weight_panel(p->w(wk), p->vals[gp].ptr, 1, cin, cout, 0, (const float *)p->vals[gv].ptr, B, st);
        });
    } else if (use_tc) {
        panel = static_panel(p, a.weight, a.k * a.k, a.cin, a.cout);
        if (!panel) return EFFDET_E_CUDA;
    }
    effdet_conv_desc *d = new effdet_conv_desc();
    memset(d, 0, sizeof(*d));
    p->descs.push_back(d);
    d->n_groups = n;
    for (int i = 0; i < n; ++i) {
        d->H[i] = a.H[i]; d->W[i] = a.Wd[i];
        d->ldc[i] = a.ldc.empty() ? 0 : a.ldc[i];
        d->y_batch_stride[i] = a.ybs.empty() ? 0 : a.ybs[i];
    }
    d->B = p->B; d->Cin = a.cin; d->Cout = a.cout; d->kh = d->kw = a.k; d->stride = a.stride;
    d->scale = a.scale; d->shift = a.shift; d->act = a.act; d->in_dtype = split ? EFFDET_BF16 : in_dt; d->out_dtype = out_dt;
    d->allow_tensor_core = use_tc ? 1 : 0;
    d->split_planes = split ? 1 : 0;
    ConvArgs args = a;
    args.xs = xs_used;
    std::vector<int> ins = xs_used, outs;
    for (int r : a.residuals) if (r >= 0) ins.push_back(r);
    if (a.gate >= 0) ins.push_back(a.gate);
    if (gate_panel >= 0) ins.push_back(gate_panel);
    for (int y : a.ys) { bool seen = false; for (int o : outs) seen |= o == y; if (!seen) outs.push_back(y); }
    add_op(p, ins, outs, [=](cudaStream_t st) {
        for (int i = 0; i < n; ++i) {
            d->x[i] = p->vals[args.xs[i]].ptr;
            d->y[i] = (char *)p->vals[args.ys[i]].ptr + (args.y_off.empty() ? 0 : args.y_off[i]);
            d->residual[i] = (args.residuals.empty() || args.residuals[i] < 0) ? nullptr : p->vals[args.residuals[i]].ptr;
        }
        d->weight = p->w(args.weight);
        d->gate = args.gate >= 0 ? (const float *)p->vals[args.gate].ptr : nullptr;
        if (gate_panel >= 0) { d->weight_bf16 = p->vals[gate_panel].ptr; d->weight_per_sample = 1; }
        else { d->weight_bf16 = panel; d->weight_per_sample = 0; }
        return effdet_conv2d(d, st);
    });
    return EFFDET_OK;
}

int emit_conv_block(effdet_plan *p, int x, int H, const std::string &name, int cin, int cout, int k, int stride,
                    int *out, int *Hout) {
    const int Ho = (H + stride - 1) / stride;
    const int y = act_val(p, Ho, Ho, cout);
    const effdet_plan::Folded &f = p->folded[p->folded_index.at(name + "_bn")];
    ConvArgs a;
    a.xs = {x}; a.ys = {y}; a.H = {H}; a.Wd = {H};
    a.weight = name + "_conv/kernel"; a.cin = cin; a.cout = cout; a.k = k; a.stride = stride;
    a.scale = f.scale; a.shift = f.shift; a.act = EFFDET_ACT_RELU;
    *out = y; *Hout = Ho;
    return emit_conv(p, a);
}

int emit_node(effdet_plan *p, int in0, int mode0, int in1, int in2, int H, const std::string &fuse, const std::string &dw,
              int *out) {
    const int W = p->W, B = p->B, dt = p->dtype;
    const int y = act_val(p, H, H, W);
    const effdet_plan::Folded &f = p->folded[p->folded_index.at(dw + "_bn")];
    const bool weighted = p->weighted != 0;
    std::vector<int> ins = {in0, in1};
    if (in2 >= 0) ins.push_back(in2);
    add_op(p, ins, {y}, [=](cudaStream_t st) {
        return effdet_bifpn_node(p->vals[in0].ptr, mode0, p->vals[in1].ptr, in2 >= 0 ? p->vals[in2].ptr : nullptr,
                                 weighted ? p->w(fuse + "/" + fuse) : nullptr, 1e-4f, p->w(dw + "_dconv/depthwise_kernel"),
                                 f.scale, f.shift, p->vals[y].ptr, B, H, H, W, dt, st);
    });
    *out = y;
    return EFFDET_OK;
}

int build_ops(effdet_plan *p) {
    const int B = p->B, S = p->S, dt = p->dtype;
    const bool u8 = (p->flags & EFFDET_PLAN_U8_INPUT) != 0;
    int rc;
    p->v_images = new_val(p, (size_t)B * S * S * 3 * (u8 ? 1 : 4), true);
    int H = (S + 1) / 2;
    int x = act_val(p, H, H, p->c0);
    {
        const effdet_plan::Folded &f = p->folded[p->folded_index.at("stem_bn")];
        const int img = p->v_images, c0 = p->c0, xo = x;
        if (u8) add_op(p, {img}, {xo}, [=](cudaStream_t st) {
            return effdet_stem_conv_u8((const unsigned char *)p->vals[img].ptr, p->lut, p->w("stem_conv/kernel"), f.scale,
                                       f.shift, p->vals[xo].ptr, B, S, S, c0, EFFDET_ACT_SWISH, dt, st);
        });
        else add_op(p, {img}, {xo}, [=](cudaStream_t st) {
            return effdet_stem_conv((const float *)p->vals[img].ptr, p->w("stem_conv/kernel"), f.scale, f.shift,
                                    p->vals[xo].ptr, B, S, S, c0, dt, st);
        });
    }
    std::vector<int> feats, featH, featC;
    for (const Block &b : p->blocks) {                       // efficientnet.py:210-306, inference form
        const std::string q = b.prefix;
        const int inp = x;
        if (b.expand != 1) {
            const int e = act_val(p, H, H, b.cmid);
            const effdet_plan::Folded &f = p->folded[p->folded_index.at(q + "expand_bn")];
            ConvArgs a;
            a.xs = {x}; a.ys = {e}; a.H = {H}; a.Wd = {H};
            a.weight = q + "expand_conv/kernel"; a.cin = b.cin; a.cout = b.cmid;
            a.scale = f.scale; a.shift = f.shift; a.act = EFFDET_ACT_SWISH;
            if ((rc = emit_conv(p, a))) return rc;
            x = e;
        }
        const int Ho = (H + b.stride - 1) / b.stride;
        // fp32 accuracy mode on the tensor cores: the depthwise kernel writes the hi | lo planes the project
        // convolution reads (engine.py _mbconv dw_split)
        static const bool fp32_tc = !(getenv("EFFDET_FP32_TC") && atoi(getenv("EFFDET_FP32_TC")) == 0);
        const bool dw_split = fp32_tc && dt == EFFDET_F32 && b.cmid % 8 == 0;
        const int dwv = dw_split ? new_val(p, (size_t)B * Ho * Ho * 2 * b.cmid * 2) : act_val(p, Ho, Ho, b.cmid);
        const int nblk = effdet_dwconv_se_blocks(B, H, H, b.cmid, b.stride, dt);
        const int part = new_val(p, (size_t)B * nblk * b.cmid * 4);
        {
            const effdet_plan::Folded &f = p->folded[p->folded_index.at(q + "bn")];
            const int xi = x, Hi = H, cm = b.cmid, k = b.k, s = b.stride;
            add_op(p, {xi}, {dwv, part}, [=](cudaStream_t st) {
                if (dw_split)
                    return effdet_dwconv_split_out((const float *)p->vals[xi].ptr, p->w(q + "dwconv/depthwise_kernel"),
                                                   f.scale, f.shift, p->vals[dwv].ptr, (float *)p->vals[part].ptr, nblk, B,
                                                   Hi, Hi, cm, k, s, EFFDET_ACT_SWISH, st);
                return effdet_dwconv(p->vals[xi].ptr, p->w(q + "dwconv/depthwise_kernel"), f.scale, f.shift, p->vals[dwv].ptr,
                                     (float *)p->vals[part].ptr, nblk, B, Hi, Hi, cm, k, s, EFFDET_ACT_SWISH, dt, st);
            });
        }
        const int gate = new_val(p, (size_t)B * b.cmid * 4);
        {
            const int cm = b.cmid, R = b.se;
            const float inv = 1.0f / (float)(Ho * Ho);
            add_op(p, {part}, {gate}, [=](cudaStream_t st) {
                return effdet_se_gate((const float *)p->vals[part].ptr, nblk, inv, p->w(q + "se_reduce/kernel"),
                                      p->w(q + "se_reduce/bias"), p->w(q + "se_expand/kernel"), p->w(q + "se_expand/bias"),
                                      (float *)p->vals[gate].ptr, B, cm, R, st);
            });
        }
        const int y = act_val(p, Ho, Ho, b.cout, b.tap);
        {
            const effdet_plan::Folded &f = p->folded[p->folded_index.at(q + "project_bn")];
            ConvArgs a;
            a.xs = {dwv}; a.ys = {y}; a.H = {Ho}; a.Wd = {Ho};
            a.weight = q + "project_conv/kernel"; a.cin = b.cmid; a.cout = b.cout;
            a.scale = f.scale; a.shift = f.shift; a.gate = gate; a.pre_split = dw_split;
            a.residuals = {b.skip ? inp : -1};               // FixedDropout is the identity at inference (:300-303)
            if ((rc = emit_conv(p, a))) return rc;
        }
        x = y; H = Ho;
        if (b.tap) { feats.push_back(y); featH.push_back(H); featC.push_back(b.cout); }
    }
    // ---- BiFPN (model.py:93-268)
    const int W = p->W;
    std::vector<int> P(5), PH(5);
    for (int i = 0; i < p->D; ++i) {
        const std::string pre = "BiFPN_" + std::to_string(i) + "_";
        int in[5], inH[5];
        if (i == 0) {
            for (int l = 0; l < 3; ++l)
                if ((rc = emit_conv_block(p, feats[2 + l], featH[2 + l], pre + "P" + std::to_string(3 + l), featC[2 + l], W, 1, 1,
                                          &in[l], &inH[l]))) return rc;
            if ((rc = emit_conv_block(p, feats[4], featH[4], pre + "P6", featC[4], W, 3, 2, &in[3], &inH[3]))) return rc;
            if ((rc = emit_conv_block(p, in[3], inH[3], pre + "P7", W, W, 3, 2, &in[4], &inH[4]))) return rc;
        } else {
            for (int l = 0; l < 5; ++l)
                if ((rc = emit_conv_block(p, P[l], PH[l], pre + "P" + std::to_string(3 + l), W, W, 1, 1, &in[l], &inH[l]))) return rc;
        }
        auto fn = [&](int j) { return fuse_name(8 * i + j); };
        const int UP = 1, DOWN = 2;
        int p6td, p5td, p4td, o3, o4, o5, o6, o7;
        if ((rc = emit_node(p, in[4], UP, in[3], -1, inH[3], fn(0), pre + "U_P6", &p6td))) return rc;
        if ((rc = emit_node(p, p6td, UP, in[2], -1, inH[2], fn(1), pre + "U_P5", &p5td))) return rc;
        if ((rc = emit_node(p, p5td, UP, in[1], -1, inH[1], fn(2), pre + "U_P4", &p4td))) return rc;
        if ((rc = emit_node(p, p4td, UP, in[0], -1, inH[0], fn(3), pre + "U_P3", &o3))) return rc;
        if ((rc = emit_node(p, o3, DOWN, p4td, in[1], inH[1], fn(4), pre + "D_P4", &o4))) return rc;
        if ((rc = emit_node(p, o4, DOWN, p5td, in[2], inH[2], fn(5), pre + "D_P5", &o5))) return rc;
        if ((rc = emit_node(p, o5, DOWN, p6td, in[3], inH[3], fn(6), pre + "D_P6", &o6))) return rc;
        if ((rc = emit_node(p, o6, DOWN, in[4], -1, inH[4], fn(7), pre + "D_P7", &o7))) return rc;
        P = {o3, o4, o5, o6, o7};
        PH = {inH[0], inH[1], inH[2], inH[3], inH[4]};
    }
    // ---- heads (model.py:271-353), all five levels per launch, outputs written in place of Concatenate (:393-398)
    const int A = 9, C = p->C;
    size_t cells = 0;
    std::vector<size_t> lvl_off(5);
    for (int l = 0; l < 5; ++l) { lvl_off[l] = cells * A; cells += (size_t)PH[l] * PH[l]; }
    p->N = cells * A;
    p->v_reg = new_val(p, (size_t)B * p->N * 4 * 4, true);
    p->v_cls = new_val(p, (size_t)B * p->N * C * 4, true);
    for (int h = 0; h < 2; ++h) {
        const std::string scope = h == 0 ? "box_head" : "class_head";
        const std::string fmt = h == 0 ? "regress_head_conv_" : "class_head_";
        const std::string fin = h == 0 ? "regress_head_conv_final" : "pyramid_classification";
        const int per = h == 0 ? 4 : C, out = h == 0 ? p->v_reg : p->v_cls;
        std::vector<int> xs = P;
        for (int i = 0; i < p->hd; ++i) {
            std::vector<int> ys(5);
            for (int l = 0; l < 5; ++l) ys[l] = act_val(p, PH[l], PH[l], W);
            const std::string n = scope + "/" + fmt + std::to_string(i);
            ConvArgs a;
            a.xs = xs; a.ys = ys; a.H = PH; a.Wd = PH;
            a.weight = n + "/kernel"; a.cin = W; a.cout = W; a.k = 3;
            a.shift = p->w(n + "/bias"); a.act = EFFDET_ACT_RELU;
            if ((rc = emit_conv(p, a))) return rc;
            xs = ys;
        }
        const std::string n = scope + "/" + fin;
        ConvArgs a;
        a.xs = xs; a.ys = std::vector<int>(5, out); a.H = PH; a.Wd = PH;
        a.weight = n + "/kernel"; a.cin = W; a.cout = A * per; a.k = 3;
        a.shift = p->w(n + "/bias"); a.act = h == 0 ? EFFDET_ACT_NONE : EFFDET_ACT_SIGMOID;
        a.ldc = std::vector<int>(5, A * per);
        a.ybs = std::vector<long long>(5, (long long)p->N * per);
        for (int l = 0; l < 5; ++l) a.y_off.push_back(lvl_off[l] * per * 4);
        a.out_dtype = EFFDET_F32;
        if ((rc = emit_conv(p, a))) return rc;
    }
    return EFFDET_OK;
}

// liveness-based buffer assignment (engine.Plan._assign_buffers): a buffer is recycled for a later value of the
// same size once its last reader has been emitted
int assign_buffers(effdet_plan *p) {
    for (size_t i = 0; i < p->ops.size(); ++i) {
        for (int v : p->ops[i].ins) p->vals[v].last_use = (int)i;
        for (int v : p->ops[i].outs) p->vals[v].last_use = (int)i;
    }
    std::map<size_t, std::vector<void *>> free_list;
    auto alloc = [&](Val &v) -> int {
        if (v.ptr) return EFFDET_OK;
        if (!v.keep) {
            auto it = free_list.find(v.bytes);
            if (it != free_list.end() && !it->second.empty()) { v.ptr = it->second.back(); it->second.pop_back(); return EFFDET_OK; }
        }
        return dev_alloc(p, &v.ptr, v.bytes);
    };
    int rc;
    for (Val &v : p->vals) if (v.last_use < 0 && (rc = alloc(v))) return rc;
    for (size_t i = 0; i < p->ops.size(); ++i) {
        for (int v : p->ops[i].outs) if ((rc = alloc(p->vals[v]))) return rc;
        for (int v : p->ops[i].ins) if ((rc = alloc(p->vals[v]))) return rc;
        std::vector<int> all = p->ops[i].ins;
        all.insert(all.end(), p->ops[i].outs.begin(), p->ops[i].outs.end());
        for (size_t a = 0; a < all.size(); ++a) {
            bool dup = false;
            for (size_t b = 0; b < a; ++b) dup |= all[b] == all[a];
            Val &v = p->vals[all[a]];
            if (!dup && v.last_use == (int)i && !v.keep) free_list[v.bytes].push_back(v.ptr);
        }
    }
    return EFFDET_OK;
}

int run_ops(effdet_plan *p, cudaStream_t st) {
    for (Op &op : p->ops) {
        const int rc = op.run(st);
        if (rc) return rc;
    }
    return EFFDET_OK;
}

// weights are all bound: fold the BatchNorms, build the static bf16 weight panels, warm up and capture the graph
int finalize(effdet_plan *p, cudaStream_t st) {
    for (size_t i = 0; i < p->bound.size(); ++i)
        if (!p->bound[i]) return fail(EFFDET_E_INVALID, "effdet_plan: weight %s has not been bound", p->winfo[i].name.c_str());
    for (auto &f : p->folded) {
        int rc = effdet_bn_fold(p->w(f.bn + "/gamma"), p->w(f.bn + "/beta"), p->w(f.bn + "/moving_mean"),
                                p->w(f.bn + "/moving_variance"), f.eps, f.scale, f.shift, f.C, st);
        if (rc) return rc;
    }
    for (auto &q : p->panels) {
        int rc = q.split ? effdet_conv_weight_panel_split(p->w(q.key), q.ptr, q.taps, q.cin, q.cout, nullptr, 0, st)
                         : effdet_conv_weight_panel(p->w(q.key), q.ptr, q.taps, q.cin, q.cout, 0, nullptr, 0, st);
        if (rc) return rc;
    }
    if (!p->graph && !(p->flags & EFFDET_PLAN_NO_GRAPH)) {
        int rc = run_ops(p, st);                              // warm-up: kernel attributes, lazy module loading
        if (rc) return rc;
        EFFDET_CUDA(cudaStreamSynchronize(st));
        if (!p->capture_stream) EFFDET_CUDA(cudaStreamCreateWithFlags(&p->capture_stream, cudaStreamNonBlocking));
        cudaGraph_t g = nullptr;
        EFFDET_CUDA(cudaStreamBeginCapture(p->capture_stream, cudaStreamCaptureModeThreadLocal));
        rc = run_ops(p, p->capture_stream);
        cudaError_t e = cudaStreamEndCapture(p->capture_stream, &g);
        if (rc) { if (g) cudaGraphDestroy(g); return rc; }
        if (e != cudaSuccess) return fail(EFFDET_E_CUDA, "effdet_plan: graph capture failed: %s", cudaGetErrorString(e));
        e = cudaGraphInstantiate(&p->graph, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) return fail(EFFDET_E_CUDA, "effdet_plan: cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
    }
    p->finalized = true;
    return EFFDET_OK;
}

int normalization_lut(effdet_plan *p) {
    // lut[c][v] = normalize_image of byte value v in channel c, each step rounded to float32 like the reference
    // (train_tpu.py:130-140, generators/common.py:418-429): float32(v) / float32(255) -> - float32(mean_c) ->
    // / float32(std_c)   (== efficientdet_b200/utils/preprocess.py normalization_lut)
    static const double mean[3] = {0.485, 0.456, 0.406}, stdv[3] = {0.229, 0.224, 0.225};
    float h[3 * 256];
    for (int c = 0; c < 3; ++c)
        for (int v = 0; v < 256; ++v) {
            volatile float x = (float)v / 255.0f;            // volatile: every step is rounded to float32
            x = x - (float)mean[c];
            x = x / (float)stdv[c];
            h[c * 256 + v] = x;
        }
    int rc = dev_alloc(p, (void **)&p->lut, sizeof(h));
    if (rc) return rc;
    if (p->host_mem) memcpy(p->lut, h, sizeof(h));           // dry run
    else EFFDET_CUDA(cudaMemcpy(p->lut, h, sizeof(h), cudaMemcpyHostToDevice));
    return EFFDET_OK;
}

int ensure_tail(effdet_plan *p, int max_det, size_t cap) {
    if (!p->anchors) {
        // utils/anchors.py:296-336 anchors_for_shape with AnchorParameters.default (:23-52): float32-rounded ratios and
        // scales promoted to double; the table is float64, the network consumes it as float32 (model.py:414-427)
        int hw[10], sizes[5] = {32, 64, 128, 256, 512}, strides[5] = {8, 16, 32, 64, 128};
        for (int l = 0; l < 5; ++l) hw[2 * l] = hw[2 * l + 1] = (p->S + (1 << (l + 3)) - 1) >> (l + 3);
        const double ratios[3] = {(double)0.5f, (double)1.0f, (double)2.0f};
        const double scales[3] = {(double)(float)pow(2.0, 0.0), (double)(float)pow(2.0, 1.0 / 3.0), (double)(float)pow(2.0, 2.0 / 3.0)};
        std::vector<double> a64(p->N * 4);
        int rc = effdet_anchors_for_shape_host(hw, sizes, strides, 5, ratios, 3, scales, 3, a64.data(), p->N);
        if (rc) return rc;
        std::vector<float> a32(a64.begin(), a64.end());
        if ((rc = dev_alloc(p, (void **)&p->anchors, a32.size() * 4))) return rc;
        EFFDET_CUDA(cudaMemcpy(p->anchors, a32.data(), a32.size() * 4, cudaMemcpyHostToDevice));
        if ((rc = dev_alloc(p, (void **)&p->boxes, (size_t)p->B * p->N * 16))) return rc;
        if ((rc = dev_alloc(p, (void **)&p->status, 16))) return rc;
    }
    if (max_det > p->out_max_det) {
        int rc;
        if ((rc = dev_alloc(p, (void **)&p->out_boxes, (size_t)p->B * max_det * 16))) return rc;
        if ((rc = dev_alloc(p, (void **)&p->out_scores, (size_t)p->B * max_det * 4))) return rc;
        if ((rc = dev_alloc(p, (void **)&p->out_labels, (size_t)p->B * max_det * 4))) return rc;
        p->out_max_det = max_det;
        p->ws_bytes = 0;
    }
    const size_t need = effdet_filter_detections_workspace_size(p->B, p->N, p->C, cap, max_det);
    if (need > p->ws_bytes || cap != p->cand_capacity) {
        if (need > p->ws_bytes) {
            int rc = dev_alloc(p, &p->ws, need + need / 4 + 256);
            if (rc) return rc;
            p->ws_bytes = need + need / 4 + 256;
        }
        p->cand_capacity = cap;
    }
    return EFFDET_OK;
}

}  // namespace

// ================================================================== C ABI
extern "C" int effdet_plan_create(int phi, int image_size, int batch, int num_classes, int weighted_bifpn, int dtype,
                                  unsigned flags, effdet_plan_t **out) {
    EFFDET_REQUIRE(out, "null output");
    EFFDET_REQUIRE(phi >= 0 && phi < 7, "phi in 0..6 (model.py:367)");
    EFFDET_REQUIRE(batch > 0 && num_classes > 0, "bad sizes");
    EFFDET_REQUIRE(image_size > 0 && image_size % 128 == 0, "image size must be a multiple of 128 (5 pyramid levels)");
    EFFDET_REQUIRE(dtype == EFFDET_F32 || dtype == EFFDET_BF16, "dtype");
    effdet_plan *p = new effdet_plan();
    p->phi = phi; p->S = image_size; p->B = batch; p->C = num_classes; p->weighted = weighted_bifpn ? 1 : 0;
    p->dtype = dtype; p->flags = flags;
    p->W = kWBifpn[phi]; p->D = 2 + phi; p->hd = 3 + phi / 3;                  // model.py:372-375
    build_manifest(p);
    p->bound.assign(p->winfo.size(), 0);
    int rc = dev_alloc(p, (void **)&p->flat, p->flat_count * 4);
    if (!rc && p->host_mem) memset(p->flat, 0, p->flat_count * 4);          // dry run
    else if (!rc) rc = cudaMemset(p->flat, 0, p->flat_count * 4) == cudaSuccess ? EFFDET_OK : EFFDET_E_CUDA;
    for (auto &f : p->folded) {
        if (!rc) rc = dev_alloc(p, (void **)&f.scale, (size_t)f.C * 4);
        if (!rc) rc = dev_alloc(p, (void **)&f.shift, (size_t)f.C * 4);
    }
    if (!rc && (flags & EFFDET_PLAN_U8_INPUT)) rc = normalization_lut(p);
    if (!rc) rc = build_ops(p);
    if (!rc) rc = assign_buffers(p);
    if (rc) { effdet_plan_destroy(p); return rc; }
    *out = p;
    return EFFDET_OK;
}

extern "C" int effdet_plan_destroy(effdet_plan_t *p) {
    if (!p) return EFFDET_OK;
    if (p->graph) cudaGraphExecDestroy(p->graph);
    if (p->capture_stream) cudaStreamDestroy(p->capture_stream);
    for (void *q : p->owned) cudaFree(q);
    for (void *q : p->owned_host) free(q);
    for (auto *d : p->descs) delete d;
    if (p->host_in) cudaFreeHost(p->host_in);
    if (p->host_out) cudaFreeHost(p->host_out);
    delete p;
    return EFFDET_OK;
}

extern "C" int effdet_plan_dry_run(effdet_plan_t *p, int *launches) {
    EFFDET_REQUIRE(p, "null plan");
    EFFDET_REQUIRE(p->host_mem, "only for a plan built without a CUDA driver under EFFDET_DRY_RUN");
    int n = 0;
    // a launch passes when it fails at its first CUDA call and nowhere earlier
    auto passed = [&](int rc) {
        ++n;
        return rc == EFFDET_OK || (rc == EFFDET_E_CUDA && strstr(effdet_last_error(), "driver version is insufficient"));
    };
    for (auto &f : p->folded) {
        const int rc = effdet_bn_fold(p->w(f.bn + "/gamma"), p->w(f.bn + "/beta"), p->w(f.bn + "/moving_mean"),
                                      p->w(f.bn + "/moving_variance"), f.eps, f.scale, f.shift, f.C, nullptr);
        if (!passed(rc)) return rc;
    }
    for (auto &q : p->panels) {
        const int rc = q.split ? effdet_conv_weight_panel_split(p->w(q.key), q.ptr, q.taps, q.cin, q.cout, nullptr, 0, nullptr)
                               : effdet_conv_weight_panel(p->w(q.key), q.ptr, q.taps, q.cin, q.cout, 0, nullptr, 0, nullptr);
        if (!passed(rc)) return rc;
    }
    for (Op &op : p->ops) {
        const int rc = op.run(nullptr);
        if (!passed(rc)) return rc;
    }
    if (launches) *launches = n;
    return EFFDET_OK;
}

extern "C" int effdet_plan_num_weights(const effdet_plan_t *p) { return p ? (int)p->winfo.size() : 0; }
extern "C" size_t effdet_plan_num_anchors(const effdet_plan_t *p) { return p ? p->N : 0; }
extern "C" int effdet_plan_num_launches(const effdet_plan_t *p) { return p ? (int)p->ops.size() : 0; }

extern "C" int effdet_plan_weight_info(const effdet_plan_t *p, int index, const char **name, int *ndim, int dims[4]) {
    EFFDET_REQUIRE(p && index >= 0 && index < (int)p->winfo.size(), "bad index");
    const WeightInfo &w = p->winfo[index];
    if (name) *name = w.name.c_str();
    if (ndim) *ndim = w.ndim;
    if (dims) for (int i = 0; i < 4; ++i) dims[i] = w.dims[i];
    return EFFDET_OK;
}

static int bind(effdet_plan_t *p, const char *const *names, const void *const *ptrs, int n, cudaMemcpyKind kind,
                cudaStream_t st) {
    EFFDET_REQUIRE(p && names && ptrs && n >= 0, "bad arguments");
    for (int i = 0; i < n; ++i) {
        std::string key = names[i] ? names[i] : "";
        if (key.size() > 2 && key.compare(key.size() - 2, 2, ":0") == 0) key.resize(key.size() - 2);   // Keras "<w>:0"
        auto it = p->windex.find(key);
        if (it == p->windex.end()) continue;                  // by_name=True: unknown names are ignored (train.py:329)
        EFFDET_REQUIRE(ptrs[i], "null weight pointer");
        const WeightInfo &w = p->winfo[it->second];
        EFFDET_CUDA(cudaMemcpyAsync(p->flat + w.offset, ptrs[i], w.count * 4, kind, st));
        p->bound[it->second] = 1;
    }
    p->finalized = false;                                     // folded BN / panels are stale
    return EFFDET_OK;
}
extern "C" int effdet_plan_bind_weights(effdet_plan_t *p, const char *const *names, const void *const *device_ptrs, int n,
                                        void *stream) {
    return bind(p, names, device_ptrs, n, cudaMemcpyDeviceToDevice, as_stream(stream));
}
extern "C" int effdet_plan_bind_weights_host(effdet_plan_t *p, const char *const *names, const void *const *host_ptrs, int n) {
    int rc = bind(p, names, host_ptrs, n, cudaMemcpyHostToDevice, nullptr);
    if (rc) return rc;
    EFFDET_CUDA(cudaStreamSynchronize(nullptr));              // pageable host memory: the caller may free it now
    return EFFDET_OK;
}

extern "C" int effdet_forward(effdet_plan_t *p, const void *images, float *regression_out, float *classification_out,
                              void *stream) {
    EFFDET_REQUIRE(p && images, "null argument");
    cudaStream_t st = as_stream(stream);
    if (!p->finalized) { int rc = finalize(p, st); if (rc) return rc; }
    if (images != p->vals[p->v_images].ptr)
        EFFDET_CUDA(cudaMemcpyAsync(p->vals[p->v_images].ptr, images, p->vals[p->v_images].bytes, cudaMemcpyDeviceToDevice, st));
    if (p->graph) EFFDET_CUDA(cudaGraphLaunch(p->graph, st));
    else { int rc = run_ops(p, st); if (rc) return rc; }
    if (regression_out) EFFDET_CUDA(cudaMemcpyAsync(regression_out, p->vals[p->v_reg].ptr, p->vals[p->v_reg].bytes, cudaMemcpyDeviceToDevice, st));
    if (classification_out) EFFDET_CUDA(cudaMemcpyAsync(classification_out, p->vals[p->v_cls].ptr, p->vals[p->v_cls].bytes, cudaMemcpyDeviceToDevice, st));
    return EFFDET_OK;
}

extern "C" int effdet_plan_buffers(effdet_plan_t *p, void **images, float **regression, float **classification) {
    EFFDET_REQUIRE(p, "null plan");
    if (images) *images = p->vals[p->v_images].ptr;
    if (regression) *regression = (float *)p->vals[p->v_reg].ptr;
    if (classification) *classification = (float *)p->vals[p->v_cls].ptr;
    return EFFDET_OK;
}

// forward + RegressBoxes + ClipBoxes + FilterDetections; the candidate workspace grows (with one device
// synchronisation to read the overflow word) until the batch's candidates fit -- NOT graph-capturable
static int detect(effdet_plan_t *p, const void *images, const float *anchors, float score_threshold, float iou_threshold,
                  int max_det, cudaStream_t st) {
    int rc = effdet_forward(p, images, nullptr, nullptr, st);
    if (rc) return rc;
    size_t worst = (size_t)p->B * p->N * p->C;
    size_t cap = p->cand_capacity ? p->cand_capacity : (size_t)p->B * 8192 > (1u << 16) ? (size_t)p->B * 8192 : (1u << 16);
    if (cap > worst) cap = worst;
    const float mean[4] = {0.f, 0.f, 0.f, 0.f}, stdv[4] = {0.2f, 0.2f, 0.2f, 0.2f};       // RegressBoxes.py:7-8
    for (;;) {
        if ((rc = ensure_tail(p, max_det, cap))) return rc;
        if ((rc = effdet_regress_clip_boxes(anchors ? anchors : p->anchors, 0, (const float *)p->vals[p->v_reg].ptr, mean, stdv, p->B,
                                            p->N, (float)p->S, (float)p->S, p->boxes, st))) return rc;
        EFFDET_CUDA(cudaMemsetAsync(p->status, 0, 16, st));
        if ((rc = effdet_filter_detections(p->boxes, (const float *)p->vals[p->v_cls].ptr, p->B, p->N, p->C, score_threshold,
                                           iou_threshold, max_det, 1, iou_threshold > 0 ? 1 : 0, p->ws, p->ws_bytes, cap,
                                           p->out_boxes, p->out_scores, p->out_labels, nullptr, p->status, st))) return rc;
        int32_t hs[4];
        EFFDET_CUDA(cudaMemcpyAsync(hs, p->status, 16, cudaMemcpyDeviceToHost, st));
        EFFDET_CUDA(cudaStreamSynchronize(st));
        if (hs[0] == 0) return EFFDET_OK;
        const size_t need = (size_t)(uint32_t)hs[1];
        cap = hs[1] >= 0x7fffffff ? worst : need + need / 8 + 1024;
        if (cap > worst) cap = worst;
    }
}

extern "C" int effdet_detect(effdet_plan_t *p, const void *images, const float *anchors, float score_threshold,
                             float iou_threshold, int max_detections, float *boxes_out, float *scores_out,
                             int32_t *labels_out, void *stream) {
    EFFDET_REQUIRE(p && images && boxes_out && scores_out && labels_out && max_detections > 0, "bad arguments");
    cudaStream_t st = as_stream(stream);
    int rc = detect(p, images, anchors, score_threshold, iou_threshold, max_detections, st);
    if (rc) return rc;
    const size_t n = (size_t)p->B * max_detections;
    EFFDET_CUDA(cudaMemcpyAsync(boxes_out, p->out_boxes, n * 16, cudaMemcpyDeviceToDevice, st));
    EFFDET_CUDA(cudaMemcpyAsync(scores_out, p->out_scores, n * 4, cudaMemcpyDeviceToDevice, st));
    EFFDET_CUDA(cudaMemcpyAsync(labels_out, p->out_labels, n * 4, cudaMemcpyDeviceToDevice, st));
    return EFFDET_OK;
}

extern "C" int effdet_detect_host(effdet_plan_t *p, const void *images_host, const float *anchors_host, float score_threshold,
                                  float iou_threshold, int max_detections, float *boxes_host, float *scores_host,
                                  int32_t *labels_host) {
    EFFDET_REQUIRE(p && images_host && boxes_host && scores_host && labels_host && max_detections > 0, "bad arguments");
    if (!p->capture_stream) EFFDET_CUDA(cudaStreamCreateWithFlags(&p->capture_stream, cudaStreamNonBlocking));
    cudaStream_t st = p->capture_stream;
    const size_t in_bytes = p->vals[p->v_images].bytes;
    if (!p->host_in) EFFDET_CUDA(cudaMallocHost(&p->host_in, in_bytes));
    memcpy(p->host_in, images_host, in_bytes);                 // caller memory may be pageable: stage through pinned
    EFFDET_CUDA(cudaMemcpyAsync(p->vals[p->v_images].ptr, p->host_in, in_bytes, cudaMemcpyHostToDevice, st));
    float *anchors = nullptr;
    if (anchors_host) {                                        // the (1,N,4) anchors input of inference.py:57-59
        if (!p->user_anchors) { int rc = dev_alloc(p, (void **)&p->user_anchors, p->N * 16); if (rc) return rc; }
        anchors = p->user_anchors;
        EFFDET_CUDA(cudaMemcpyAsync(anchors, anchors_host, p->N * 16, cudaMemcpyHostToDevice, st));
    }
    int rc = detect(p, p->vals[p->v_images].ptr, anchors, score_threshold, iou_threshold, max_detections, st);
    if (rc) return rc;
    const size_t n = (size_t)p->B * max_detections;
    if (p->host_out_bytes < n * 24) {
        if (p->host_out) cudaFreeHost(p->host_out);
        EFFDET_CUDA(cudaMallocHost(&p->host_out, n * 24));
        p->host_out_bytes = n * 24;
    }
    char *ho = (char *)p->host_out;
    EFFDET_CUDA(cudaMemcpyAsync(ho, p->out_boxes, n * 16, cudaMemcpyDeviceToHost, st));
    EFFDET_CUDA(cudaMemcpyAsync(ho + n * 16, p->out_scores, n * 4, cudaMemcpyDeviceToHost, st));
    EFFDET_CUDA(cudaMemcpyAsync(ho + n * 20, p->out_labels, n * 4, cudaMemcpyDeviceToHost, st));
    EFFDET_CUDA(cudaStreamSynchronize(st));
    memcpy(boxes_host, ho, n * 16);
    memcpy(scores_host, ho + n * 16, n * 4);
    memcpy(labels_host, ho + n * 20, n * 4);
    return EFFDET_OK;
}
