// Compiled training plans (the training half of the plan-level C ABI): replays, without Python, the
// launch list of one optimizer step that efficientdet_b200/plan_export.py wrote -- device target assignment,
// forward in training mode, focal + smooth-L1 losses, backward, SGD (train_tpu.py:249-346 on MirroredStrategy's
// per-replica batch).  The file holds, per launch, the name of one of THIS library's entry points and its
// arguments with device pointers as (region, offset), plus the contents of the state regions (weights, optimizer
// velocity, static operands).  effdet_replay_load allocates the regions, relocates the pointers and captures the
// launches into one CUDA graph; effdet_replay_step replays it with the step's learning rate.
// Host code only: no kernels live in this file; the launches go through the generated thunks (replay_thunks.inc).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "common.cuh"
#include "effdet_b200.h"

namespace {

union ReplayArg { void *p; long long i; double d; };
struct ReplayThunk { const char *name; int n_args; int (*fn)(const ReplayArg *, void *); };
#include "replay_thunks.inc"

enum { A_I64 = 0, A_F64 = 1, A_PTR = 2, A_NULL = 3, A_BLOB = 4, A_LR = 5 };

struct Region { std::string name; size_t bytes = 0; void *dev = nullptr; };
struct Launch {
    int lane = 0;
    std::vector<uint32_t> waits;                    // launches on other lanes this one follows
    const ReplayThunk *thunk = nullptr;
    std::vector<ReplayArg> args;
    std::vector<int> lr_args;                       // argument slots that take the step's learning rate
    std::vector<std::vector<uint8_t>> blobs;        // host structs / arrays the arguments point to (relocated)
};

struct Reader {
    FILE *f;
    bool ok = true;
    void get(void *dst, size_t n) { if (ok && fread(dst, 1, n, f) != n) ok = false; }
    template <typename T> T v() { T t{}; get(&t, sizeof(T)); return t; }
    std::string str(size_t n) { std::string s(n, '\0'); if (n) get(&s[0], n); return s; }
};

}  // namespace

struct effdet_replay {
    std::vector<Region> regions;
    std::vector<Launch> launches;
    int n_lanes = 1;
    cudaStream_t lane_streams[8] = {};              // lanes 1.. (lane 0 is the capturing stream)
    std::vector<cudaEvent_t> events;                // one per launch (capture-time dependencies)
    cudaEvent_t start_event = nullptr;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    float *lr_dev = nullptr;                        // the "lr" region (absent in plans with a by-value rate)
    float lr_host[64] = {};
    int lr_slot = 0;
    long long steps = 0;
    bool use_graph = true;
};

using effdet::fail;

// Eager: plan order on one stream.  Capturing: the plan's multi-lane schedule -- lanes are streams forked from the
// capturing stream, cross-lane dependencies are event edges, all lanes join the capturing stream at the end (the
// last launch, the optimizer, waits for their tails).
static int run_launches(effdet_replay *p, double lr, cudaStream_t st, bool lanes) {
    if (!lanes || p->n_lanes <= 1) {
        for (Launch &l : p->launches) {
            for (int k : l.lr_args) l.args[k].d = lr;
            const int rc = l.thunk->fn(l.args.data(), st);
            if (rc != EFFDET_OK) return rc;
        }
        return EFFDET_OK;
    }
    if (p->events.empty()) {
        p->events.resize(p->launches.size());
        for (cudaEvent_t &e : p->events) EFFDET_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        EFFDET_CUDA(cudaEventCreateWithFlags(&p->start_event, cudaEventDisableTiming));
        for (int l = 1; l < p->n_lanes; ++l) EFFDET_CUDA(cudaStreamCreateWithFlags(&p->lane_streams[l], cudaStreamNonBlocking));
    }
    EFFDET_CUDA(cudaEventRecord(p->start_event, st));
    bool joined[8] = {true, false, false, false, false, false, false, false};
    for (size_t i = 0; i < p->launches.size(); ++i) {
        Launch &l = p->launches[i];
        cudaStream_t ls = l.lane == 0 ? st : p->lane_streams[l.lane];
        if (!joined[l.lane]) { EFFDET_CUDA(cudaStreamWaitEvent(ls, p->start_event, 0)); joined[l.lane] = true; }
        for (uint32_t j : l.waits) EFFDET_CUDA(cudaStreamWaitEvent(ls, p->events[j], 0));
        for (int k : l.lr_args) l.args[k].d = lr;
        const int rc = l.thunk->fn(l.args.data(), ls);
        if (rc != EFFDET_OK) return rc;
        EFFDET_CUDA(cudaEventRecord(p->events[i], ls));
    }
    return EFFDET_OK;
}

extern "C" int effdet_replay_destroy(effdet_replay_t *p) {
    if (!p) return EFFDET_OK;
    if (p->exec) cudaGraphExecDestroy(p->exec);
    if (p->graph) cudaGraphDestroy(p->graph);
    for (cudaEvent_t e : p->events) if (e) cudaEventDestroy(e);
    if (p->start_event) cudaEventDestroy(p->start_event);
    for (int l = 1; l < 8; ++l) if (p->lane_streams[l]) cudaStreamDestroy(p->lane_streams[l]);
    for (Region &r : p->regions)
        if (r.dev) cudaFree(r.dev);
    delete p;
    return EFFDET_OK;
}

extern "C" int effdet_replay_load(const char *path, int flags, effdet_replay_t **out) {
    EFFDET_REQUIRE(path && out, "null argument");
    *out = nullptr;
    FILE *f = fopen(path, "rb");
    if (!f) return fail(EFFDET_E_INVALID, "effdet_replay_load: cannot open %s (%lld)", path, 0LL);
    Reader rd{f};
    effdet_replay *p = new effdet_replay();
    p->use_graph = !(flags & 1);
    auto bail = [&](const char *why) {
        fclose(f);
        effdet_replay_destroy(p);
        return fail(EFFDET_E_INVALID, "effdet_replay_load: %s (%lld)", why, 0LL);
    };
    char magic[8];
    rd.get(magic, 8);
    if (!rd.ok || memcmp(magic, "EFDPLAN1", 8) != 0) return bail("not a compiled plan");
    const uint32_t n_regions = rd.v<uint32_t>();
    if (!rd.ok || n_regions > (1u << 20)) return bail("bad region count");
    p->regions.resize(n_regions);
    std::vector<uint8_t> host;
    for (Region &r : p->regions) {
        r.bytes = (size_t)rd.v<uint64_t>();
        const uint8_t has = rd.v<uint8_t>();
        r.name = rd.str(rd.v<uint16_t>());
        if (!rd.ok) return bail("truncated region table");
        if (cudaMalloc(&r.dev, r.bytes ? r.bytes : 16) != cudaSuccess) return bail("cudaMalloc failed");
        if (has) {
            host.resize(r.bytes);
            rd.get(host.data(), r.bytes);
            if (!rd.ok) return bail("truncated region contents");
            if (cudaMemcpy(r.dev, host.data(), r.bytes, cudaMemcpyHostToDevice) != cudaSuccess) return bail("upload failed");
        } else if (cudaMemset(r.dev, 0, r.bytes) != cudaSuccess) {
            return bail("cudaMemset failed");
        }
    }
    auto resolve = [&](uint32_t region, uint64_t off, void **ptr) {
        if (region >= p->regions.size() || off > p->regions[region].bytes) return false;
        *ptr = static_cast<uint8_t *>(p->regions[region].dev) + off;
        return true;
    };
    const uint32_t n_ops = rd.v<uint32_t>();
    if (!rd.ok || n_ops > (1u << 24)) return bail("bad launch count");
    p->launches.resize(n_ops);
    for (Launch &l : p->launches) {
        l.lane = rd.v<uint8_t>();
        const int n_waits = rd.v<uint16_t>();
        for (int q = 0; q < n_waits; ++q) l.waits.push_back(rd.v<uint32_t>());
        if (!rd.ok || l.lane >= 8) return bail("bad lane");
        if (l.lane + 1 > p->n_lanes) p->n_lanes = l.lane + 1;
        for (uint32_t j : l.waits)
            if (j >= (uint32_t)(&l - p->launches.data())) return bail("bad dependency");
        const std::string name = rd.str(rd.v<uint16_t>());
        const int n_args = rd.v<uint16_t>();
        if (!rd.ok) return bail("truncated launch list");
        for (const ReplayThunk &t : kReplayThunks)
            if (name == t.name) l.thunk = &t;
        if (!l.thunk) return bail(("unknown entry point " + name).c_str());
        if (l.thunk->n_args != n_args) return bail(("argument count mismatch for " + name).c_str());
        l.args.resize(n_args);
        l.blobs.reserve(n_args);
        for (int k = 0; k < n_args; ++k) {
            ReplayArg &a = l.args[k];
            a.i = 0;
            switch (rd.v<uint8_t>()) {
            case A_I64: a.i = rd.v<int64_t>(); break;
            case A_F64: a.d = rd.v<double>(); break;
            case A_NULL: a.p = nullptr; break;
            case A_LR: a.d = 0.0; l.lr_args.push_back(k); break;
            case A_PTR: {
                const uint32_t region = rd.v<uint32_t>();
                const uint64_t off = rd.v<uint64_t>();
                if (!rd.ok || !resolve(region, off, &a.p)) return bail("bad device pointer");
                break;
            }
            case A_BLOB: {
                const uint32_t len = rd.v<uint32_t>();
                if (!rd.ok || len > (1u << 20)) return bail("bad blob");
                l.blobs.emplace_back(len);
                std::vector<uint8_t> &b = l.blobs.back();
                rd.get(b.data(), len);
                const int n_rel = rd.v<uint16_t>();
                for (int q = 0; q < n_rel; ++q) {
                    const uint32_t at = rd.v<uint32_t>(), region = rd.v<uint32_t>();
                    const uint64_t off = rd.v<uint64_t>();
                    void *ptr = nullptr;
                    if (!rd.ok || at + sizeof(void *) > len || !resolve(region, off, &ptr)) return bail("bad relocation");
                    memcpy(b.data() + at, &ptr, sizeof(void *));
                }
                a.p = b.data();
                break;
            }
            default: return bail("bad argument kind");
            }
        }
        if (!rd.ok) return bail("truncated launch");
    }
    fclose(f);
    for (Region &r : p->regions)
        if (r.name == "lr") p->lr_dev = static_cast<float *>(r.dev);
    *out = p;
    return EFFDET_OK;
}

extern "C" int effdet_replay_num_launches(const effdet_replay_t *p) { return p ? (int)p->launches.size() : 0; }

extern "C" int effdet_replay_region(effdet_replay_t *p, const char *name, void **device_ptr, size_t *bytes) {
    EFFDET_REQUIRE(p && name, "null argument");
    for (Region &r : p->regions)
        if (r.name == name) {
            if (device_ptr) *device_ptr = r.dev;
            if (bytes) *bytes = r.bytes;
            return EFFDET_OK;
        }
    return fail(EFFDET_E_INVALID, "effdet_replay_region: no region named %s (%lld)", name, 0LL);
}

/* One optimizer step on the batch in the "images" / "gt_*" regions.  The learning rate is written into the "lr"
 * region (the SGD launch reads it from there), then the launches run: captured into a CUDA graph on the second call
 * (the first one runs eagerly: kernel attributes are set on first use) and replayed from then on. */
extern "C" int effdet_replay_step(effdet_replay_t *p, double learning_rate, void *stream) {
    EFFDET_REQUIRE(p, "null plan");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (p->lr_dev) {
        p->lr_host[p->lr_slot] = (float)learning_rate;      // rotating slots: the copy is asynchronous
        EFFDET_CUDA(cudaMemcpyAsync(p->lr_dev, &p->lr_host[p->lr_slot], sizeof(float), cudaMemcpyHostToDevice, st));
        p->lr_slot = (p->lr_slot + 1) % 64;
    }
    if (!p->use_graph || p->steps == 0) {
        ++p->steps;
        return run_launches(p, learning_rate, st, false);
    }
    if (!p->exec) {
        cudaStream_t own = nullptr;
        EFFDET_CUDA(cudaStreamCreateWithFlags(&own, cudaStreamNonBlocking));
        EFFDET_CUDA(cudaStreamBeginCapture(own, cudaStreamCaptureModeThreadLocal));
        const int rc = run_launches(p, learning_rate, own, true);
        cudaGraph_t g = nullptr;
        const cudaError_t e = cudaStreamEndCapture(own, &g);
        cudaStreamDestroy(own);
        if (rc != EFFDET_OK) { if (g) cudaGraphDestroy(g); return rc; }
        EFFDET_CUDA(e);
        p->graph = g;
        EFFDET_CUDA(cudaGraphInstantiate(&p->exec, p->graph, 0));
    }
    ++p->steps;
    EFFDET_CUDA(cudaGraphLaunch(p->exec, st));
    return EFFDET_OK;
}
