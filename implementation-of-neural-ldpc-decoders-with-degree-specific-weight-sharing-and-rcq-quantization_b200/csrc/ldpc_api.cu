// C-ABI layer (include/ldpc_b200.h): Tanner-graph re-layout, decoder objects, device workspace,
// the per-iteration launch sequence, the chunked host<->device pipeline and the Monte-Carlo round.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/ldpc_b200.h"
#include "ldpc_internal.h"

using namespace ldpc;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(e_ == cudaErrorMemoryAllocation ? LDPC_ERR_NOMEM : LDPC_ERR_CUDA,         \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; }
        if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

template <typename T>
int upload(T** dptr, const std::vector<T>& h) {
    *dptr = nullptr;
    size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
    CU(cudaMalloc((void**)dptr, bytes));
    if (!h.empty()) CU(cudaMemcpy(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return LDPC_OK;
}

constexpr int kCnChunk = 8;   // checks per work item
constexpr int kVnChunk = 8;   // variables per work item
constexpr int64_t kFrameAlign = 128;

}  // namespace

// =================================================================================================
// Graph
// =================================================================================================
struct ldpc_graph {
    int device = 0;
    int32_t n = 0, m = 0;
    int64_t E = 0;
    int max_dc = 0, max_dv = 0;
    int n_cclass = 0, n_vclass = 0;
    // host copies (layout + tests)
    std::vector<int32_t> slot_of_edge;   // [E] check-major edge id -> slot
    std::vector<int32_t> slot_var;       // [E] slot -> variable
    std::vector<int32_t> vslots;         // slot lists by degree-sorted variable position
    std::vector<int32_t> vpos_var;       // [n] position -> variable
    std::vector<int32_t> var_vpos;       // [n] variable -> position
    // work items: [0] coarse (8 nodes per item, large batches), [1] fine (1 node per item: small batches are
    // latency-bound, so the serial chain inside an item is what a half iteration takes)
    struct ItemList {
        std::vector<WorkItem> items;
        WorkItem* d = nullptr;
        int wide_begin = 0, wide_end = 0, wide_max_deg = 0;   // items [begin, end): degree 9..64
    };
    ItemList cn[2], vn[2];
    std::vector<int64_t> chk_ptr;        // original CSR (layered schedule walks checks in index order)
    std::vector<int32_t> chk_var;
    int nonempty_checks = 0;
    // dependency levels of the layered schedule (checks in index order; see layered_level_kernel)
    std::vector<int32_t> level_ptr, level_chk;
    int32_t* d_level_chk = nullptr;
    // software-pipelined sequential walk (layered_pipe_kernel): one record per non-empty check; empty when a
    // check has more than kLayerMaxDeg edges
    std::vector<LayerRec> lay_recs;
    LayerRec* d_lay_recs = nullptr;
    // CTA-resident decode (ldpc_resident.cu): degree classes and 16-bit index tables in the kernel's own slot order
    // ("physical" slots: the checks of a class in tiles of 32, edge k of check (tile, lane) at
    // first_slot + tile * 32 * deg + k * 32 + lane, so that the lanes of a warp -- consecutive checks -- touch
    // consecutive shared-memory words at compile-time strides).  Built when E, n < 65536.
    struct Resident {
        bool ok = false;
        int n_cclass = 0, n_vclass = 0;
        int32_t E_phys = 0;                  // physical slots (classes padded to whole 32-node tiles)
        std::vector<int32_t> phys;           // [E] slot -> physical slot
        WorkItem* d_classes = nullptr;       // check classes, then variable classes
        uint16_t* d_slot_var = nullptr;      // [E] physical slot -> variable
        uint16_t* d_vslots = nullptr;        // per variable class: entry d of its i-th variable at first_slot + d * count + i
        uint16_t* d_vpos_var = nullptr;      // [n] position -> variable
    } res;
    // device copies
    int64_t* d_chk_ptr = nullptr;
    int32_t* d_chk_var = nullptr;
    int32_t* d_slot_var = nullptr;
    int32_t* d_vslots = nullptr;
    int32_t* d_vpos_var = nullptr;
};

extern "C" int ldpc_version(void) { return LDPC_B200_VERSION; }
extern "C" const char* ldpc_last_error(void) { return g_err.c_str(); }

extern "C" int ldpc_device_count(int* count) {
    if (!count) return fail(LDPC_ERR_INVALID, "count is NULL");
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) {
        *count = 0;
        return fail(LDPC_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    *count = c;
    return LDPC_OK;
}

extern "C" int ldpc_host_alloc(void** ptr, int64_t bytes) {
    if (!ptr || bytes < 0) return fail(LDPC_ERR_INVALID, "bad arguments");
    CU(cudaHostAlloc(ptr, (size_t)std::max<int64_t>(bytes, 1), cudaHostAllocDefault));
    return LDPC_OK;
}
extern "C" int ldpc_host_free(void* ptr) {
    if (ptr) CU(cudaFreeHost(ptr));
    return LDPC_OK;
}

// K6 graph_relayout: replaces the dense-H neighbour scans (ldpc_decoder.py:85,92,124,136).
extern "C" int ldpc_graph_create(int device, int32_t n, int32_t m, const int64_t* check_ptr,
                                 const int32_t* check_var, ldpc_graph** out) {
    if (!out) return fail(LDPC_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (n <= 0 || m < 0 || !check_ptr || (m > 0 && check_ptr[m] > 0 && !check_var))
        return fail(LDPC_ERR_INVALID, "bad graph arguments");
    if (check_ptr[0] != 0) return fail(LDPC_ERR_INVALID, "check_ptr[0] must be 0");
    const int64_t E = check_ptr[m];
    if (E >= (int64_t)1 << 31) return fail(LDPC_ERR_UNSUPPORTED, "more than 2^31-1 edges");
    std::vector<int32_t> dv(n, 0);
    int max_dc = 0;
    for (int32_t i = 0; i < m; ++i) {
        int64_t a = check_ptr[i], b = check_ptr[i + 1];
        if (b < a) return fail(LDPC_ERR_INVALID, "check_ptr not monotone at %d", i);
        max_dc = std::max<int64_t>(max_dc, b - a);
        for (int64_t e = a; e < b; ++e) {
            int32_t j = check_var[e];
            if (j < 0 || j >= n) return fail(LDPC_ERR_INVALID, "variable index %d out of range", j);
            if (e > a && check_var[e - 1] >= j)
                return fail(LDPC_ERR_INVALID, "check %d: variables must be strictly ascending", i);
            dv[j]++;
        }
    }
    int max_dv = 0;
    for (int32_t j = 0; j < n; ++j) max_dv = std::max(max_dv, dv[j]);

    ldpc_graph* g = new (std::nothrow) ldpc_graph();
    if (!g) return fail(LDPC_ERR_NOMEM, "host allocation failed");
    g->device = device;
    g->n = n;
    g->m = m;
    g->E = E;
    g->max_dc = max_dc;
    g->max_dv = max_dv;
    g->chk_ptr.assign(check_ptr, check_ptr + m + 1);
    g->chk_var.assign(check_var, check_var + E);
    for (int32_t i = 0; i < m; ++i) g->nonempty_checks += check_ptr[i + 1] > check_ptr[i];

    // ---- check side: stable sort of non-empty checks by degree -> slots ----
    std::vector<int32_t> corder;
    corder.reserve(m);
    for (int32_t i = 0; i < m; ++i)
        if (check_ptr[i + 1] > check_ptr[i]) corder.push_back(i);
    std::stable_sort(corder.begin(), corder.end(), [&](int32_t a, int32_t b) {
        return (check_ptr[a + 1] - check_ptr[a]) < (check_ptr[b + 1] - check_ptr[b]);
    });
    g->slot_of_edge.assign((size_t)E, -1);
    g->slot_var.assign((size_t)E, 0);
    {
        int32_t slot = 0;
        int prev_deg = -1;
        size_t pos = 0;
        while (pos < corder.size()) {
            int deg = (int)(check_ptr[corder[pos] + 1] - check_ptr[corder[pos]]);
            if (deg != prev_deg) {
                g->n_cclass++;
                prev_deg = deg;
            }
            size_t end = pos;
            while (end < corder.size() && (int)(check_ptr[corder[end] + 1] - check_ptr[corder[end]]) == deg) ++end;
            for (size_t c = pos; c < end; c += kCnChunk) {
                WorkItem it;
                it.deg = deg;
                it.count = (int32_t)std::min<size_t>(kCnChunk, end - c);
                it.first_node = (int32_t)c;
                it.first_slot = slot + (int32_t)((c - pos) * (size_t)deg);
                g->cn[0].items.push_back(it);
            }
            for (size_t c = pos; c < end; ++c) {
                int32_t i = corder[c];
                for (int k = 0; k < deg; ++k) {
                    int64_t e = check_ptr[i] + k;
                    g->slot_of_edge[(size_t)e] = slot;
                    g->slot_var[(size_t)slot] = check_var[e];
                    ++slot;
                }
            }
            pos = end;
        }
    }
    // ---- variable side: stable sort by degree; slot lists in ascending check index ----
    std::vector<int32_t> vorder(n);
    for (int32_t j = 0; j < n; ++j) vorder[j] = j;
    std::stable_sort(vorder.begin(), vorder.end(), [&](int32_t a, int32_t b) { return dv[a] < dv[b]; });
    g->vpos_var = vorder;
    g->var_vpos.assign(n, 0);
    std::vector<int64_t> lbase(n, 0);
    {
        int64_t off = 0;
        for (int32_t p = 0; p < n; ++p) {
            g->var_vpos[vorder[p]] = p;
            lbase[vorder[p]] = off;
            off += dv[vorder[p]];
        }
    }
    g->vslots.assign((size_t)E, 0);
    {
        std::vector<int32_t> fill(n, 0);
        for (int32_t i = 0; i < m; ++i)  // ascending check index => ascending inside every variable list
            for (int64_t e = check_ptr[i]; e < check_ptr[i + 1]; ++e) {
                int32_t j = check_var[e];
                g->vslots[(size_t)(lbase[j] + fill[j]++)] = g->slot_of_edge[(size_t)e];
            }
        int prev_deg = -1;
        int32_t pos = 0;
        while (pos < n) {
            int deg = dv[vorder[pos]];
            if (deg != prev_deg) {
                g->n_vclass++;
                prev_deg = deg;
            }
            int32_t end = pos;
            while (end < n && dv[vorder[end]] == deg) ++end;
            for (int32_t c = pos; c < end; c += kVnChunk) {
                WorkItem it;
                it.deg = deg;
                it.count = std::min<int32_t>(kVnChunk, end - c);
                it.first_node = c;
                it.first_slot = (int32_t)lbase[vorder[c]];
                g->vn[0].items.push_back(it);
            }
            pos = end;
        }
    }
    // ---- dependency levels of the checks in index order (layered schedule) ----
    {
        std::vector<int32_t> last(n, 0), lvl(m, 0);
        int32_t nlev = 0;
        for (int32_t i = 0; i < m; ++i) {
            if (check_ptr[i + 1] == check_ptr[i]) continue;
            int32_t l = 0;
            for (int64_t e = check_ptr[i]; e < check_ptr[i + 1]; ++e) l = std::max(l, last[check_var[e]]);
            lvl[i] = ++l;
            for (int64_t e = check_ptr[i]; e < check_ptr[i + 1]; ++e) last[check_var[e]] = l;
            nlev = std::max(nlev, l);
        }
        g->level_ptr.assign((size_t)nlev + 1, 0);
        for (int32_t i = 0; i < m; ++i)
            if (lvl[i]) g->level_ptr[(size_t)lvl[i]]++;
        for (int32_t l = 0; l < nlev; ++l) g->level_ptr[(size_t)l + 1] += g->level_ptr[(size_t)l];
        g->level_chk.assign((size_t)g->nonempty_checks, 0);
        std::vector<int32_t> fill(g->level_ptr.begin(), g->level_ptr.end());
        for (int32_t i = 0; i < m; ++i)   // ascending check index inside a level
            if (lvl[i]) g->level_chk[(size_t)fill[(size_t)lvl[i] - 1]++] = i;
    }
    // ---- records of the software-pipelined sequential walk (layered_pipe_kernel) ----
    if (max_dc <= kLayerMaxDeg) {
        const int depth = layered_pipe_depth();
        std::vector<LayerRec> recs;
        recs.reserve((size_t)g->nonempty_checks);
        std::vector<int32_t> last_s(n, -1), last_k(n, 0);   // latest check of the walk that touched a variable, and where
        for (int32_t i = 0; i < m; ++i) {
            const int64_t e0 = check_ptr[i], e1 = check_ptr[i + 1];
            if (e1 == e0) continue;
            const int32_t s = (int32_t)recs.size();
            LayerRec r{};
            r.dc = (uint8_t)(e1 - e0);
            for (int64_t e = e0; e < e1; ++e) {
                const int k = (int)(e - e0);
                const int32_t v = check_var[e];
                r.var[k] = v;
                if (last_s[v] >= 0 && s - last_s[v] < depth)
                    recs[(size_t)last_s[v]].desc[last_k[v]] |= (uint8_t)(((s - last_s[v]) << 3) | k);   // forwarded by its last writer
                else
                    r.ahead_mask |= (uint8_t)(1u << k);                                                 // copied ahead
                last_s[v] = s;
                last_k[v] = k;
            }
            recs.push_back(r);
        }
        g->lay_recs = std::move(recs);
    }
    // fine lists (one node per item) and the degree ranges of both
    for (ldpc_graph::ItemList* pair : {g->cn, g->vn}) {
        for (const WorkItem& it : pair[0].items)
            for (int c = 0; c < it.count; ++c) {
                WorkItem f = it;
                f.count = 1;
                f.first_node = it.first_node + c;
                f.first_slot = it.first_slot + c * it.deg;
                pair[1].items.push_back(f);
            }
        for (int k = 0; k < 2; ++k) {
            ldpc_graph::ItemList& L = pair[k];
            const int nitems = (int)L.items.size();
            L.wide_begin = L.wide_end = nitems;
            for (int i = 0; i < nitems; ++i) {
                const int deg = L.items[(size_t)i].deg;
                if (deg > 8 && L.wide_begin == nitems) L.wide_begin = i;
                if (deg > 64) { L.wide_end = i; break; }
                if (deg > 8) L.wide_max_deg = deg;
            }
            if (L.wide_end < L.wide_begin) L.wide_begin = L.wide_end;   // no degree in 9..64
        }
    }
    // ---- tables of the CTA-resident decode ----
    std::vector<WorkItem> res_classes;
    std::vector<uint16_t> res_slot_var, res_vslots, res_vpos_var;
    if (E > 0 && E < 65536 && n < 65536) {
        auto classes_of = [](const std::vector<WorkItem>& fine) {
            std::vector<WorkItem> cls;
            for (const WorkItem& it : fine) {
                if (!cls.empty() && cls.back().deg == it.deg) cls.back().count++;
                else cls.push_back(it);   // count == 1, first_node / first_slot of the class's first node
            }
            return cls;
        };
        std::vector<WorkItem> cc = classes_of(g->cn[1].items), vc = classes_of(g->vn[1].items);
        if ((int)cc.size() <= kResMaxClasses && (int)vc.size() <= kResMaxClasses) {
            // tiles of 32 nodes: inside a tile, edge k of node (lane) at tile_base + k * 32 + lane -- compile-time strides
            // for the kernel's unrolled node code; a class is padded to whole tiles
            auto tiled = [](int32_t base, int node, int k, int deg) { return base + (node >> 5) * (32 * deg) + k * 32 + (node & 31); };
            ldpc_graph::Resident& r = g->res;
            r.ok = true;
            r.n_cclass = (int)cc.size();
            r.n_vclass = (int)vc.size();
            r.phys.assign((size_t)E, 0);
            int32_t pbase = 0;
            for (WorkItem& cl : cc) {
                for (int c = 0; c < cl.count; ++c)
                    for (int k = 0; k < cl.deg; ++k) r.phys[(size_t)(cl.first_slot + c * cl.deg + k)] = tiled(pbase, c, k, cl.deg);
                cl.first_slot = pbase;
                pbase += (cl.count + 31) / 32 * 32 * cl.deg;
            }
            r.E_phys = pbase;
            res_slot_var.assign((size_t)r.E_phys, 0);
            for (int64_t sl = 0; sl < E; ++sl) res_slot_var[(size_t)r.phys[(size_t)sl]] = (uint16_t)g->slot_var[(size_t)sl];
            int32_t vbase = 0;
            std::vector<std::pair<size_t, uint16_t>> entries;
            entries.reserve((size_t)E);
            for (WorkItem& cl : vc) {
                for (int i = 0; i < cl.count; ++i)
                    for (int dd = 0; dd < cl.deg; ++dd)
                        entries.emplace_back((size_t)tiled(vbase, i, dd, cl.deg),
                                             (uint16_t)r.phys[(size_t)g->vslots[(size_t)(cl.first_slot + i * cl.deg + dd)]]);
                cl.first_slot = vbase;
                vbase += (cl.count + 31) / 32 * 32 * cl.deg;
            }
            res_vslots.assign((size_t)std::max<int32_t>(vbase, 1), 0);
            for (const auto& e : entries) res_vslots[e.first] = e.second;
            res_vpos_var.assign(g->vpos_var.begin(), g->vpos_var.end());
            res_classes = cc;
            res_classes.insert(res_classes.end(), vc.begin(), vc.end());
            if (r.E_phys >= 65536) r.ok = false;   // physical slots are 16-bit table entries
        }
    }
    // ---- upload ----
    DeviceGuard guard(device);
    if (!guard.ok) {
        delete g;
        return fail(LDPC_ERR_CUDA, "cannot select CUDA device %d", device);
    }
    int rc = upload(&g->d_slot_var, g->slot_var);
    if (!rc) rc = upload(&g->d_vslots, g->vslots);
    if (!rc) rc = upload(&g->d_vpos_var, g->vpos_var);
    for (int k = 0; k < 2 && !rc; ++k) {
        rc = upload(&g->cn[k].d, g->cn[k].items);
        if (!rc) rc = upload(&g->vn[k].d, g->vn[k].items);
    }
    if (!rc) rc = upload(&g->d_chk_ptr, g->chk_ptr);
    if (!rc) rc = upload(&g->d_chk_var, g->chk_var);
    if (!rc) rc = upload(&g->d_level_chk, g->level_chk);
    if (!rc) rc = upload(&g->d_lay_recs, g->lay_recs);
    if (!rc && g->res.ok) {
        rc = upload(&g->res.d_classes, res_classes);
        if (!rc) rc = upload(&g->res.d_slot_var, res_slot_var);
        if (!rc) rc = upload(&g->res.d_vslots, res_vslots);
        if (!rc) rc = upload(&g->res.d_vpos_var, res_vpos_var);
    }
    if (rc) {
        ldpc_graph_destroy(g);
        return rc;
    }
    *out = g;
    return LDPC_OK;
}

extern "C" int ldpc_graph_destroy(ldpc_graph* g) {
    if (!g) return LDPC_OK;
    DeviceGuard guard(g->device);
    cudaFree(g->d_slot_var);
    cudaFree(g->d_vslots);
    cudaFree(g->d_vpos_var);
    for (int k = 0; k < 2; ++k) {
        cudaFree(g->cn[k].d);
        cudaFree(g->vn[k].d);
    }
    cudaFree(g->d_chk_ptr);
    cudaFree(g->d_chk_var);
    cudaFree(g->d_level_chk);
    cudaFree(g->d_lay_recs);
    cudaFree(g->res.d_classes);
    cudaFree(g->res.d_slot_var);
    cudaFree(g->res.d_vslots);
    cudaFree(g->res.d_vpos_var);
    delete g;
    return LDPC_OK;
}

extern "C" int ldpc_graph_query(const ldpc_graph* g, int what, int64_t* value) {
    if (!g || !value) return fail(LDPC_ERR_INVALID, "NULL argument");
    switch (what) {
        case LDPC_GRAPH_N: *value = g->n; break;
        case LDPC_GRAPH_M: *value = g->m; break;
        case LDPC_GRAPH_E: *value = g->E; break;
        case LDPC_GRAPH_CHECK_CLASSES: *value = g->n_cclass; break;
        case LDPC_GRAPH_VAR_CLASSES: *value = g->n_vclass; break;
        case LDPC_GRAPH_MAX_DC: *value = g->max_dc; break;
        case LDPC_GRAPH_MAX_DV: *value = g->max_dv; break;
        case LDPC_GRAPH_DEVICE: *value = g->device; break;
        case LDPC_GRAPH_LAYER_LEVELS: *value = (int64_t)g->level_ptr.size() - 1; break;
        case LDPC_GRAPH_LAYER_PIPED: *value = g->lay_recs.empty() ? 0 : 1; break;
        default: return fail(LDPC_ERR_INVALID, "unknown query %d", what);
    }
    return LDPC_OK;
}

extern "C" int ldpc_graph_slot_of_edge(const ldpc_graph* g, int32_t* slot_of_edge) {
    if (!g || !slot_of_edge) return fail(LDPC_ERR_INVALID, "NULL argument");
    std::memcpy(slot_of_edge, g->slot_of_edge.data(), (size_t)g->E * sizeof(int32_t));
    return LDPC_OK;
}

// =================================================================================================
// Decoder
// =================================================================================================
namespace {

struct Workspace {
    int64_t cap = 0;  // frames (multiple of kFrameAlign)
    void* llrT = nullptr;
    void* v2c = nullptr;
    void* c2v = nullptr;
    uint32_t* hardw = nullptr;
    uint32_t* unsat = nullptr;  // [2][Wn]
    uint8_t* done = nullptr;
    int32_t* iters = nullptr;
    uint8_t* success = nullptr;
    int32_t* fcnt = nullptr;    // [Bp] per-frame bit-error scratch of the Monte-Carlo count
    void* post = nullptr;       // [n][cap] posterior rows; allocated the first time a posterior is asked for
    void release() {
        cudaFree(fcnt);
        cudaFree(post);
        cudaFree(llrT); cudaFree(v2c); cudaFree(c2v); cudaFree(hardw); cudaFree(unsat);
        cudaFree(done); cudaFree(iters); cudaFree(success);
        *this = Workspace();
    }
};

constexpr int kPipeDepth = 4;   // staging buffers: two chunks decoding, one arriving, one leaving
constexpr size_t kMaxLevels = 32;   // compaction levels (each at most 60 % of its parent)

// Host-buffer pipeline: four streams (H2D copies, two kernel streams, D2H copies) over kPipeDepth staging
// buffers; the copy of chunk i+2, the decodes of chunks i and i+1 (two contexts of the decoder) and the
// copy-out of chunk i-1 overlap.
struct HostPipe {
    cudaStream_t s_in = nullptr, s_run = nullptr, s_run2 = nullptr, s_out = nullptr;
    struct Buf {
        void* d_llr = nullptr;
        uint8_t* d_bits = nullptr;
        void* d_post = nullptr;
        int32_t* d_it = nullptr;
        uint8_t* d_su = nullptr;
        cudaEvent_t in_ready = nullptr, run_done = nullptr, out_done = nullptr;
    } buf[kPipeDepth];
    int64_t cap = 0;
    bool post_cap = false;
    void release() {
        for (auto& b : buf) {
            cudaFree(b.d_llr); cudaFree(b.d_bits); cudaFree(b.d_post); cudaFree(b.d_it); cudaFree(b.d_su);
            if (b.in_ready) cudaEventDestroy(b.in_ready);
            if (b.run_done) cudaEventDestroy(b.run_done);
            if (b.out_done) cudaEventDestroy(b.out_done);
            b = Buf();
        }
        if (s_in) cudaStreamDestroy(s_in);
        if (s_run) cudaStreamDestroy(s_run);
        if (s_run2) cudaStreamDestroy(s_run2);
        if (s_out) cudaStreamDestroy(s_out);
        s_in = s_run = s_run2 = s_out = nullptr;
        cap = 0;
        post_cap = false;
    }
};

}  // namespace

struct ldpc_decoder {
    ldpc_graph* g = nullptr;
    int dtype = LDPC_F32;
    int V = 4;
    size_t rsz = 4;
    int T = 0;
    int early_stop = 1;
    int n_beta = 0, n_alpha = 0, bc = 0, Q = 0, nth = 0;
    std::vector<int> mono;                 // per quantiser
    std::vector<int32_t> q_of_iter;        // host copy
    int32_t* d_bidx = nullptr;             // per slot
    int beta_per_edge = 0;                 // some check mixes beta columns
    int32_t* d_aidx = nullptr;             // per vpos
    int32_t* d_aidx_slot = nullptr;        // per slot (offset rule: alpha is applied at the check node)
    int32_t* d_res_bidx = nullptr;         // d_bidx / d_aidx_slot in the slot order of the CTA-resident decode
    int32_t* d_res_aidx_slot = nullptr;
    int check_rule = 0, schedule = 0;
    int wide_ring = 1;                     // checks of degree 9..64 through the bulk-async row ring
    void* d_beta = nullptr;                // [T][n_beta]
    void* d_alpha = nullptr;               // [T][n_alpha]
    float* d_thr = nullptr;                // [Q][nth]
    float* d_lut = nullptr;                // [Q][2^bc]
    int32_t* d_qoi = nullptr;              // [T] quantiser of each iteration (on-chip decode)
    int32_t* d_mono = nullptr;             // [Q]
    int all_mono = 1;
    int use_small = 1;                     // LDPC_SMALL=0: never take the on-chip decode for small codes
    int64_t stat_small = 0;
    int use_resident = -1;                 // LDPC_RESIDENT: 1 always / 0 never take the CTA-resident decode; -1 = policy
    int sm_count = 0;
    int64_t stat_resident = 0;
    int64_t resident_max_frames = -1;      // LDPC_RESIDENT_MAX_FRAMES: batches up to this many frames decode CTA-resident (-1: a few waves)
    int64_t resident_wave_frames = -1;     // frames one wave of the resident kernel holds (occupancy query, first use)
    HostPipe pipe;
    int64_t host_chunk = 0;
    int layered_levels = 1;       // LDPC_LAYERED_LEVELS=0: always the sequential layered kernel
    int layered_stage = 1;        // LDPC_LAYERED_STAGE=0: level-parallel kernel without the shared-memory stage (A/B)
    int64_t layered_v2_frames = 65536;   // two frames per thread in the pipelined walk from this batch size on (LDPC_LAYERED_V2_FRAMES, 0 = never)
    int64_t layered_v4_frames = 0;       // four frames per thread from this batch size on (LDPC_LAYERED_V4_FRAMES, 0 = never: measured
                                         // 1012 / 1169 / 1115 K frames/s at 65 536 / 131 072 / 262 144 frames against 1023 / 1103 / 1153 K)
    int layered_pipe = 1;         // LDPC_LAYERED_PIPE=0: the plain sequential kernel instead of the software-pipelined one
    int host_dual = 1;            // LDPC_HOST_DUAL=0: one chunk decodes at a time in the host pipeline
    // frame compaction (early stop at scale): child workspaces, one per level, plus bookkeeping buffers
    struct Level {
        Workspace ws;
        int32_t* idx = nullptr;   // [cap] running frames of the parent, ascending
        int32_t* map = nullptr;   // [cap] frame of this level -> frame of the caller's batch
        int64_t cap = 0;
    };
    // Everything ONE in-flight decode mutates.  cx[0] serves the device-pointer entry points and the Monte-Carlo
    // round, cx[1] / cx[2] the two chunks the host pipeline keeps in flight (their kernels fill each other's tails).
    struct Ctx {
        Workspace root;
        std::vector<Level> levels;    // capacity reserved at creation: jobs keep pointers into it
        int32_t* d_scan = nullptr;    // [scan_cap] per-block counts / offsets
        int64_t scan_cap = 0;
        int32_t* d_total = nullptr;   // device int32
        // two checkpoints may be outstanding (the host decides one span late): slot = checkpoint number & 1
        int32_t* h_total = nullptr;            // mapped pinned host int32[2], written by the scan kernel itself: a
                                               // copy on the D2H engine would queue behind the output copies of the
                                               // host pipeline (hundreds of MB with posteriors) and stall the job
        int32_t* h_total_dev = nullptr;        // the same words as the device sees them
        cudaEvent_t ev[2] = {nullptr, nullptr};   // recorded after a checkpoint's count has been copied to h_total[slot]
        // end of the last call that used this context, and the stream it ran on: a call on ANOTHER stream first
        // waits for it (the workspace is shared), and so does a weight update
        cudaEvent_t tail = nullptr;
        cudaStream_t tail_stream = nullptr;
        bool tail_valid = false;
    };
    Ctx cx[3];
    // launch-bound decodes (tiny codes / tiny batches): the whole T-iteration launch sequence is captured once
    // into a CUDA graph per (workspace, B, posterior) and replayed with one launch
    struct GraphEntry {
        const void* llrT = nullptr;   // identifies the workspace allocation the nodes point into
        int64_t B = 0, Bp = 0;
        bool want_post = false;
        cudaGraphExec_t exec = nullptr;
        int64_t launches = 0, cn_launches = 0, vn_launches = 0;
        uint64_t stamp = 0;
    };
    std::vector<GraphEntry> graphs;
    cudaStream_t cap_stream = nullptr;
    uint64_t graph_clock = 0;
    int use_graphs = 1;                 // LDPC_GRAPHS=0 switches the replay path off
    int64_t graph_max_iter_bytes = (int64_t)128 << 20;   // "launch-bound": one iteration moves less than this
    int64_t stat_graph_replays = 0;
    int64_t fine_items_max_frames = 128;    // batches up to this size (single-frame calls) use one-node work items
    int compact = 1;              // LDPC_COMPACT=0 switches compaction and the all-done exit off
    int64_t compact_min_frames = 512;
    int compact_percent = 60;     // compact when at most this share of the level's lanes still runs
    int checkpoint_step = 1;      // iterations between checkpoints for batches of >= 16384 frames
    int speculate = 0;            // LDPC_SPECULATE: 1 always / -1 while no frame has stopped yet / 0 (default) never keep a second
                                  // span in flight while the host waits for a checkpoint (see job_top_up)
    int post_mode = 0;            // LDPC_POST_MODE: 0 adaptive, 1 posterior rows refreshed every iteration, 2 written on stop
    int train_general = 0;        // LDPC_TRAIN_GENERAL=1: every check takes the general four-pass backward form (A/B of the three-pass one)
    int64_t stat_compactions = 0, stat_early_exits = 0;
    // posterior training: the forward pass keeps every iteration's messages for the backward pass
    struct TrainCtx {
        Workspace ws;
        float* v2c_hist = nullptr;   // [T][E][cap]
        float* c2v_hist = nullptr;   // [T][E][cap]
        float* g_v2c = nullptr;      // [E][cap]
        float* g_c2v = nullptr;      // [E][cap]
        float* g_post = nullptr;     // [n][cap]
        float* g_part = nullptr;     // spread copies of the weight gradients (ldpc_train_backward)
        size_t part_cap = 0;         // floats
        int64_t cap = 0;
        int64_t B = 0, Bp = 0;       // of the last forward pass (0: none yet)
    } train;
    // instrumentation
    int prof_mode = 0;
    ldpc_profile prof{};
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool;
    struct Pending { int kind; cudaEvent_t a, b; };
    std::vector<Pending> pending;
    size_t ev_next = 0;
};

namespace {

int64_t pad_frames(int64_t B) { return (B + kFrameAlign - 1) / kFrameAlign * kFrameAlign; }

// which work-item list a launch on Bp frames uses: fine (1) for small batches, coarse (0) otherwise
int item_set(const ldpc_decoder* d, int64_t Bp) { return Bp <= d->fine_items_max_frames ? 1 : 0; }

int ws_ensure(ldpc_decoder* d, Workspace& ws, int64_t Bp) {
    if (ws.cap >= Bp) return LDPC_OK;
    if (Bp > ((int64_t)1 << 28)) return fail(LDPC_ERR_UNSUPPORTED, "more than 2^28 frames per call (32-bit row strides)");
    for (auto& ge : d->graphs)   // captured graphs point into workspace allocations: drop them all
        if (ge.exec) cudaGraphExecDestroy(ge.exec);
    d->graphs.clear();
    ws.release();
    const ldpc_graph* g = d->g;
    // the layered schedule works in place on the posteriors (llrT): it has no message arrays
    const bool layered = d->schedule == LDPC_SCHEDULE_LAYERED;
    const size_t rows_v2c = layered ? 0 : (size_t)std::max<int64_t>(g->E, 1);
    const size_t rows_c2v = layered ? 0 : (size_t)std::max<int64_t>(g->E, 1);
    const size_t c2v_elt = d->bc ? 1 : d->rsz;
    const int64_t Wn = Bp / 32;
    CU(cudaMalloc(&ws.llrT, (size_t)g->n * Bp * d->rsz));
    CU(cudaMalloc(&ws.v2c, std::max<size_t>(rows_v2c * Bp * d->rsz, 16)));
    CU(cudaMalloc(&ws.c2v, std::max<size_t>(rows_c2v * Bp * c2v_elt, 16)));
    CU(cudaMalloc((void**)&ws.hardw, (size_t)g->n * Wn * sizeof(uint32_t)));
    CU(cudaMalloc((void**)&ws.unsat, (size_t)2 * Wn * sizeof(uint32_t)));
    CU(cudaMalloc((void**)&ws.done, (size_t)Bp));
    CU(cudaMalloc((void**)&ws.iters, (size_t)Bp * sizeof(int32_t)));
    CU(cudaMalloc((void**)&ws.success, (size_t)Bp));
    CU(cudaMalloc((void**)&ws.fcnt, (size_t)Bp * sizeof(int32_t)));
    ws.cap = Bp;
    return LDPC_OK;
}

// forward()'s posterior rows, [n][cap] like llrT (only decoders that are asked for posteriors pay for them)
int post_ensure(ldpc_decoder* d, Workspace& ws) {
    if (ws.post) return LDPC_OK;
    CU(cudaMalloc(&ws.post, (size_t)d->g->n * (size_t)ws.cap * d->rsz));
    return LDPC_OK;
}

// ---- instrumentation helpers ----
enum { K_CN = 0, K_VN = 1, K_OTHER = 2 };

struct Timed {
    ldpc_decoder* d;
    int kind;
    cudaStream_t s;
    cudaEvent_t a = nullptr, b = nullptr;
    Timed(ldpc_decoder* d_, int kind_, cudaStream_t s_) : d(d_), kind(kind_), s(s_) {
        d->prof.launches++;
        if (kind == K_CN) d->prof.cn_launches++;
        if (kind == K_VN) d->prof.vn_launches++;
        if (d->prof_mode == 1) {
            if (d->ev_next >= d->ev_pool.size()) {
                cudaEvent_t x, y;
                if (cudaEventCreate(&x) == cudaSuccess && cudaEventCreate(&y) == cudaSuccess)
                    d->ev_pool.emplace_back(x, y);
            }
            if (d->ev_next < d->ev_pool.size()) {
                a = d->ev_pool[d->ev_next].first;
                b = d->ev_pool[d->ev_next].second;
                d->ev_next++;
                cudaEventRecord(a, s);
            }
        }
    }
    ~Timed() {
        if (a) {
            cudaEventRecord(b, s);
            d->pending.push_back({kind, a, b});
        }
    }
};

#define LAUNCH(kind, expr)                                                                     \
    do {                                                                                       \
        cudaError_t le_;                                                                       \
        {                                                                                      \
            Timed t_(d, kind, stream);                                                         \
            le_ = (expr);                                                                      \
        }                                                                                      \
        if (le_ != cudaSuccess)                                                                \
            return fail(LDPC_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(le_));              \
    } while (0)

// The flooding schedule on frames already resident as llrT [n][Bp] in `ws`.
// After it returns (stream order): ws.hardw holds the final hard decisions of every frame,
// ws.iters / ws.success the per-frame results, and ws.post holds postT [n][Bp] if want_post.
// Layered RCQ (rcq_decoder.py:281-350): the posteriors live in ws.llrT and are updated in place.
int run_layered(ldpc_decoder* d, Workspace& ws, int64_t B, int64_t Bp, bool want_post, cudaStream_t stream) {
    const ldpc_graph* g = d->g;
    const int64_t Wn = Bp / 32;
    LAUNCH(K_OTHER, launch_reset_state(ws.done, ws.iters, ws.success, ws.unsat, B, Bp, d->T, stream));
    for (int t = 0; t < d->T; ++t) {
        const int q = d->q_of_iter[t];
        const int nlev = (int)g->level_ptr.size() - 1;
        if (d->layered_levels && nlev > 0 && (int64_t)nlev * 16 <= g->nonempty_checks) {
            // few dependency levels (e.g. one per block row of a quasi-cyclic code): a level's checks run concurrently
            for (int l = 0; l < nlev; ++l)
                LAUNCH(K_CN, launch_layered_level(static_cast<float*>(ws.llrT), g->d_chk_ptr, g->d_chk_var,
                                                  g->d_level_chk + g->level_ptr[(size_t)l],
                                                  g->level_ptr[(size_t)l + 1] - g->level_ptr[(size_t)l],
                                                  d->d_thr + (size_t)q * d->nth, d->nth, d->mono[q], ws.done, Bp, g->max_dc,
                                                  (d->layered_stage && Bp >= 16384) ? d->layered_stage : 0, stream));   // (8192 frames: 955 K staged, 1035 K not)
        } else if (d->layered_pipe && !g->lay_recs.empty()) {
            // chain-structured codes: one thread per frame walks the checks, inputs prefetched / forwarded on chip
            LAUNCH(K_CN, launch_layered_pipe(static_cast<float*>(ws.llrT), g->d_lay_recs, (int)g->lay_recs.size(),
                                             d->d_thr + (size_t)q * d->nth, d->nth, d->mono[q], ws.done, Bp,
                                             (d->layered_v4_frames > 0 && Bp >= d->layered_v4_frames) ? 4
                                             : (d->layered_v2_frames > 0 && Bp >= d->layered_v2_frames) ? 2 : 1, stream));
        } else {
            LAUNCH(K_CN, launch_layered_iter(static_cast<float*>(ws.llrT), g->d_chk_ptr, g->d_chk_var, g->m,
                                             d->d_thr + (size_t)q * d->nth, d->nth, d->bc, d->mono[q], ws.done, Bp, stream));
        }
        LAUNCH(K_VN, launch_hard(d->dtype, ws.llrT, ws.hardw, Wn, g->n, Bp, stream));
        if (d->early_stop || t == d->T - 1) {
            uint32_t* cur = ws.unsat + (size_t)(t & 1) * Wn;
            uint32_t* nxt = ws.unsat + (size_t)((t + 1) & 1) * Wn;
            SynLaunch sy{};
            sy.hardw = ws.hardw;
            sy.Wn = Wn;
            sy.slot_var = g->d_slot_var;
            sy.items = g->cn[item_set(d, Bp)].d;
            sy.n_items = (int)g->cn[item_set(d, Bp)].items.size();
            sy.unsat = cur;
            LAUNCH(K_OTHER, launch_syndrome(sy, stream));
            LAUNCH(K_OTHER, launch_commit(d->V, cur, nxt, ws.done, ws.iters, ws.success, t + 1, Bp, stream));
        }
    }
    if (want_post) CU(cudaMemcpyAsync(ws.post, ws.llrT, (size_t)g->n * Bp * d->rsz, cudaMemcpyDeviceToDevice, stream));
    return LDPC_OK;
}

// Where the results of a decode go: row-major user buffers (decode) or error counters (Monte-Carlo round).
struct OutSpec {
    uint8_t* bits = nullptr;
    uint32_t* packed = nullptr;      // decisions as packed rows [B][ceil(n/32)] instead of / besides one byte per bit
    void* post = nullptr;
    int32_t* iters = nullptr;
    uint8_t* success = nullptr;
    bool count = false;              // Monte-Carlo: accumulate into counters instead
    const uint8_t* codeword = nullptr;
    int64_t* counters = nullptr;
    int32_t* frame_bit_errors = nullptr;
    int32_t* frame_iters = nullptr;
};

void fill_cn(ldpc_decoder* d, Workspace& ws, int64_t Bp, int t, CnLaunch& cn) {
    const ldpc_graph* g = d->g;
    const int q = d->bc ? d->q_of_iter[t] : 0;
    cn.src = (t == 0) ? ws.llrT : ws.v2c;
    cn.dst = ws.c2v;
    cn.row_map = (t == 0) ? g->d_slot_var : nullptr;
    cn.bidx = d->d_bidx;
    cn.beta_per_edge = d->beta_per_edge;
    cn.beta_t = d->d_beta ? (const char*)d->d_beta + (size_t)t * d->n_beta * d->rsz : nullptr;
    cn.thr = d->bc ? d->d_thr + (size_t)q * d->nth : nullptr;
    cn.nth = d->nth;
    cn.bc = d->bc;
    cn.mono = d->bc ? d->mono[q] : 1;
    cn.done = ws.done;
    const ldpc_graph::ItemList& il = g->cn[item_set(d, Bp)];
    cn.items = il.d;
    cn.n_items = (int)il.items.size();
    cn.items_wide_begin = il.wide_begin;
    cn.items_wide_end = il.wide_end;
    cn.wide_ring = d->wide_ring;
    cn.Bp = Bp;
    if (d->check_rule == LDPC_RULE_OFFSET) {
        cn.aidx_slot = d->d_aidx_slot;
        cn.alpha_t = d->d_alpha ? (const char*)d->d_alpha + (size_t)t * d->n_alpha * d->rsz : nullptr;
    }
}

void fill_vn(ldpc_decoder* d, Workspace& ws, int64_t Bp, int t, bool final_pass, bool want_post, VnLaunch& vn) {
    const ldpc_graph* g = d->g;
    vn.iters = ws.iters;
    vn.post_iter = 0;
    vn.c2v = ws.c2v;
    vn.v2c = ws.v2c;
    vn.llrT = ws.llrT;
    vn.postT = want_post ? ws.post : nullptr;   // running frames refresh their posterior every iteration
    vn.vslots = g->d_vslots;
    vn.vpos_var = g->d_vpos_var;
    vn.aidx = d->d_aidx;
    vn.alpha_t = (d->d_alpha && d->check_rule != LDPC_RULE_OFFSET)
                     ? (const char*)d->d_alpha + (size_t)t * d->n_alpha * d->rsz : nullptr;
    vn.lut = d->d_lut;
    vn.bc = d->bc;
    vn.n_quant = d->Q;
    vn.q_now = d->bc ? d->q_of_iter[t] : 0;
    vn.hardw = ws.hardw;
    vn.Wn = Bp / 32;
    vn.done = ws.done;
    const ldpc_graph::ItemList& il = g->vn[item_set(d, Bp)];
    vn.items = il.d;
    vn.n_items = (int)il.items.size();
    vn.Bp = Bp;
    vn.final_pass = final_pass ? 1 : 0;
    vn.items_wide_begin = il.wide_begin;
    vn.items_wide_end = il.wide_end;
    vn.wide_max_deg = il.wide_max_deg;
    vn.wide_stage = d->wide_ring;
}

// How forward()'s posterior rows ws.post are kept.  Invariant after every iteration of either mode: a frame that
// has stopped holds the posterior of the iteration it stopped at.
//   POST_EACH:    every running frame refreshes its entries in every variable-node pass (+4n bytes written per
//                 frame-iteration, and as many read where a lane mixes stopped and running frames).
//   POST_ON_STOP: nothing is written until a frame stops; right after the commit of iteration t a posterior-only
//                 pass recomputes the posterior of the frames that stopped AT t from the check->variable messages
//                 of t (still in place: the next check-node pass has not run).  Iteration T-1 writes the rest.
//                 Costs nothing while no frame stops, but a lane (4 frames) / DRAM sector (8 frames) with one newly
//                 stopped frame reads all its message rows again, so it loses once a few percent of the frames stop
//                 per iteration.
enum { POST_NONE = 0, POST_EACH = 1, POST_ON_STOP = 2 };

// Flooding iterations [t0, t1) on the frames of `ws`.  Iteration T-1 runs the FINAL variable-node variant (its
// dead v2c update is not written).  Stopped frames are never touched again: their packed decisions stay in
// place from the iteration they stopped at, and so does their posterior row entry when posteriors are wanted.
int run_span(ldpc_decoder* d, Workspace& ws, int64_t Bp, int t0, int t1, int post_mode, cudaStream_t stream) {
    const ldpc_graph* g = d->g;
    const int64_t Wn = Bp / 32;
    for (int t = t0; t < t1; ++t) {
        CnLaunch cn{};
        fill_cn(d, ws, Bp, t, cn);
        if (d->check_rule == LDPC_RULE_OFFSET) LAUNCH(K_CN, launch_cn_offset(d->dtype, cn, stream));
        else LAUNCH(K_CN, launch_cn(d->dtype, cn, stream));
        const bool last = (t == d->T - 1);
        VnLaunch vn{};
        fill_vn(d, ws, Bp, t, last, post_mode == POST_EACH || (post_mode == POST_ON_STOP && last), vn);
        LAUNCH(K_VN, launch_vn(d->dtype, vn, stream));
        if (d->early_stop || last) {
            uint32_t* cur = ws.unsat + (size_t)(t & 1) * Wn;
            uint32_t* nxt = ws.unsat + (size_t)((t + 1) & 1) * Wn;
            SynLaunch sy{};
            sy.hardw = ws.hardw;
            sy.Wn = Wn;
            sy.slot_var = g->d_slot_var;
            sy.items = g->cn[item_set(d, Bp)].d;
            sy.n_items = (int)g->cn[item_set(d, Bp)].items.size();
            sy.unsat = cur;
            LAUNCH(K_OTHER, launch_syndrome(sy, stream));
            LAUNCH(K_OTHER, launch_commit(d->V, cur, nxt, ws.done, ws.iters, ws.success, t + 1, Bp, stream));
            if (post_mode == POST_ON_STOP && !last) {
                VnLaunch pv{};
                fill_vn(d, ws, Bp, t, true, true, pv);
                pv.post_iter = t + 1;
                LAUNCH(K_OTHER, launch_vn(d->dtype, pv, stream));
            }
        }
    }
    return LDPC_OK;
}

// Results of one level -> the caller's buffers.  `map` (nullptr = identity) gives the caller's frame of each
// frame of the level; `only_done` != nullptr leaves the running frames to the next level.
int emit_level(ldpc_decoder* d, Workspace& ws, int64_t B, int64_t Bp, const int32_t* map, const uint8_t* only_done,
               const OutSpec& o, cudaStream_t stream) {
    const ldpc_graph* g = d->g;
    if (o.count) {
        LAUNCH(K_OTHER, launch_count_packed(d->V, ws.hardw, Bp / 32, g->n, B, o.codeword, ws.iters, o.counters,
                                            o.frame_bit_errors, o.frame_iters, map, only_done, ws.fcnt, stream));
        return LDPC_OK;
    }
    // running frames of a parent level also write their (unfinished) rows here; the next level overwrites them
    if (o.bits) LAUNCH(K_OTHER, launch_unpack_bits(d->V, ws.hardw, Bp / 32, o.bits, B, g->n, map, stream));
    if (o.packed) LAUNCH(K_OTHER, launch_pack_rows(d->V, ws.hardw, Bp / 32, o.packed, B, g->n, map, stream));
    if (o.post) LAUNCH(K_OTHER, launch_unpack_post(d->dtype, ws.post, o.post, B, Bp, g->n, map, stream));
    if (map) {
        if (o.iters || o.success)
            LAUNCH(K_OTHER, launch_scatter_frames(ws.iters, ws.success, o.iters, o.success, map, B, stream));
    } else if (o.iters || o.success) {
        // (a kernel, not cudaMemcpyAsync: copies of the run streams must not queue on a copy engine behind the host
        // pipeline's bulk transfers)
        LAUNCH(K_OTHER, launch_scatter_frames(ws.iters, ws.success, o.iters, o.success, nullptr, B, stream));
    }
    return LDPC_OK;
}

// On-chip decode (ldpc_small.cu) of the frames resident in `ws`, if the code is small enough for it.
bool fill_small(ldpc_decoder* d, Workspace* ws, int64_t B, int64_t Bp, bool want_post, SmallLaunch& sp) {
    const ldpc_graph* g = d->g;
    if (!d->use_small || d->schedule != LDPC_SCHEDULE_FLOODING) return false;
    if (ws) {
        sp.llrT = ws->llrT;
        sp.postT = want_post ? ws->post : nullptr;
        sp.hardw = ws->hardw;
        sp.done = ws->done;
        sp.iters = ws->iters;
        sp.success = ws->success;
    }
    sp.Wn = Bp / 32;
    sp.B = B;
    sp.Bp = Bp;
    sp.T = d->T;
    sp.early_stop = d->early_stop;
    sp.n = g->n;
    sp.E = (int)g->E;
    sp.n_checks = (int)g->cn[1].items.size();
    sp.max_dv = g->max_dv;
    sp.cn_items = g->cn[1].d;
    sp.vn_items = g->vn[1].d;
    sp.slot_var = g->d_slot_var;
    sp.vslots = g->d_vslots;
    sp.vpos_var = g->d_vpos_var;
    sp.bidx = d->d_bidx;
    sp.beta_per_edge = d->beta_per_edge;
    sp.beta = d->d_beta;
    sp.n_beta = d->n_beta;
    sp.aidx = d->d_aidx;
    sp.aidx_slot = d->d_aidx_slot;
    sp.alpha = d->d_alpha;
    sp.n_alpha = d->n_alpha;
    sp.check_rule = d->check_rule;
    sp.bc = d->bc;
    sp.nth = d->nth;
    sp.n_quant = d->Q;
    sp.thr = d->d_thr;
    sp.lut = d->d_lut;
    sp.q_of_iter = d->d_qoi;
    sp.mono = d->d_mono;
    sp.all_mono = d->all_mono;
    return small_decode_fits(d->dtype, sp);
}

// CTA-resident decode (ldpc_resident.cu), if a frame's messages fit one SM's shared memory.
bool fill_resident(ldpc_decoder* d, int64_t B, ResidentLaunch& rp) {
    const ldpc_graph* g = d->g;
    if (d->use_resident == 0 || d->dtype != LDPC_F32 || d->schedule != LDPC_SCHEDULE_FLOODING || !g->res.ok) return false;
    rp.B = B;
    rp.T = d->T;
    rp.early_stop = d->early_stop;
    rp.n = g->n;
    rp.E = g->res.E_phys;
    rp.max_dv = g->max_dv;
    rp.n_cclass = g->res.n_cclass;
    rp.n_vclass = g->res.n_vclass;
    rp.classes = g->res.d_classes;
    rp.slot_var = g->res.d_slot_var;
    rp.vslots = g->res.d_vslots;
    rp.vpos_var = g->res.d_vpos_var;
    rp.bidx = d->d_res_bidx;
    rp.beta_per_edge = d->beta_per_edge;
    rp.beta = static_cast<const float*>(d->d_beta);
    rp.n_beta = d->n_beta;
    rp.aidx = d->d_aidx;
    rp.aidx_slot = d->d_res_aidx_slot;
    rp.alpha = static_cast<const float*>(d->d_alpha);
    rp.n_alpha = d->n_alpha;
    rp.check_rule = d->check_rule;
    rp.bc = d->bc;
    rp.nth = d->nth;
    rp.n_quant = d->Q;
    rp.thr = d->d_thr;
    rp.lut = d->d_lut;
    rp.q_of_iter = d->d_qoi;
    rp.mono = d->d_mono;
    rp.all_mono = d->all_mono;
    rp.sm_count = d->sm_count > 0 ? d->sm_count : 148;
    if (!resident_decode_fits(rp)) return false;
    if (d->use_resident == 1) return true;
    // policy (tools/resident_latency_probe.py, profiles/r02w_resident_latency.jsonl): the resident decode is issue-bound
    // (~400 G edge-iterations/s with every SM busy) where the per-iteration kernels are HBM-bound (16 bytes per edge and
    // iteration: ~400 G edge-iterations/s as well, 650 G with RCQ codes) -- but those need lanes-over-frames batches of
    // thousands of frames to fill the machine and ~4 launches per iteration, so a batch of a few WAVES (one thread block
    // per frame, as many blocks as the device holds) is faster resident: (16200,7200)-shaped 1 / 148 / 592 frames
    // 225 / 231 / 908 us against 292 / 645 / 1238 us, QC shape 148 frames 150 against 1011 us, n = 504 1184 frames 50
    // against 318 us; break-even at 4-8 waves with one block per SM, ~2 waves where an SM holds many small frames.
    if (d->resident_wave_frames < 0) d->resident_wave_frames = resident_wave_frames(rp);
    const int64_t waves = d->resident_wave_frames <= rp.sm_count ? 4 : 2;
    const int64_t max_frames = d->resident_max_frames >= 0 ? d->resident_max_frames : waves * d->resident_wave_frames;
    return B <= max_frames;
}

// Iterations after which the number of running frames is read back (one 4-byte copy + stream sync each).
// Large batches can afford a look after every iteration (an iteration is milliseconds long); smaller ones
// space the checkpoints out so that the syncs stay a small part of the decode.  `quiet` counts the
// checkpoints in a row at which no frame had stopped yet: until the first frame stops the interval is doubled.
int next_checkpoint(int t, int T, int64_t Bp, int quiet, int big_step) {
    int step;
    if (Bp >= 16384) step = t < 20 ? big_step : 2 * big_step;   // an iteration is milliseconds long: look every time
    else if (Bp >= 4096) step = t < 20 ? 2 : 4;
    else if (t < 8) step = 2;
    else if (t < 24) step = 4;
    else step = 8;
    if (quiet > 0) step *= 2;   // one doubling only: the first frames to stop must not be noticed much later
    const int c = t + step;
    return c < T ? c : T;
}

int scan_ensure(ldpc_decoder::Ctx& cx, int64_t Bp) {
    const int64_t nb = (Bp + 1023) / 1024;
    if (cx.scan_cap < nb) {
        cudaFree(cx.d_scan);
        cx.d_scan = nullptr;
        cx.scan_cap = 0;
        CU(cudaMalloc((void**)&cx.d_scan, (size_t)nb * sizeof(int32_t)));
        cx.scan_cap = nb;
    }
    if (!cx.d_total) CU(cudaMalloc((void**)&cx.d_total, sizeof(int32_t)));
    if (!cx.h_total) {
        CU(cudaHostAlloc((void**)&cx.h_total, 2 * sizeof(int32_t), cudaHostAllocMapped));
        CU(cudaHostGetDevicePointer((void**)&cx.h_total_dev, cx.h_total, 0));
    }
    for (auto& e : cx.ev)
        if (!e) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    return LDPC_OK;
}

// A decode is launch-bound when one iteration moves so few bytes (tiny code or tiny batch) that the ~4 kernel
// launches per iteration cost more than their work, and the remaining iterations of an all-stopped batch are
// cheaper to replay than a host round trip per checkpoint.
bool launch_bound(const ldpc_decoder* d, int64_t Bp) {
    if (!d->use_graphs || d->prof_mode != 0 || d->T > 24) return false;
    const ldpc_graph* g = d->g;
    const int64_t per_frame = (int64_t)(4 * g->E + g->n) * (int64_t)d->rsz;   // 16E + 4n bytes for float32
    return Bp * per_frame <= d->graph_max_iter_bytes;
}

// reset + iterations 0..T-1 as ONE graph launch.  The graph is captured on an internal stream (the caller's may
// be the legacy default stream, which cannot be captured) and its nodes point into `ws`, so an entry is keyed by
// the workspace allocation, B (pad frames are born done), Bp and the posterior flag.  Early stop still works
// inside the replay (done masks; warps of stopped frames exit at once), there are just no checkpoints.
int replay_graph(ldpc_decoder* d, Workspace& ws, int64_t B, int64_t Bp, bool want_post, cudaStream_t stream) {
    ldpc_decoder::GraphEntry* hit = nullptr;
    for (auto& ge : d->graphs)
        if (ge.exec && ge.llrT == ws.llrT && ge.B == B && ge.Bp == Bp && ge.want_post == want_post) hit = &ge;
    if (!hit) {
        if (!d->cap_stream) CU(cudaStreamCreateWithFlags(&d->cap_stream, cudaStreamNonBlocking));
        if (d->graphs.size() >= 8) {   // evict the least recently used entry
            size_t lru = 0;
            for (size_t i = 1; i < d->graphs.size(); ++i)
                if (d->graphs[i].stamp < d->graphs[lru].stamp) lru = i;
            if (d->graphs[lru].exec) cudaGraphExecDestroy(d->graphs[lru].exec);
            d->graphs.erase(d->graphs.begin() + (long)lru);
        }
        const ldpc_profile before = d->prof;
        cudaStream_t cs = d->cap_stream;
        CU(cudaStreamBeginCapture(cs, cudaStreamCaptureModeRelaxed));
        int rc = LDPC_OK;
        {
            cudaStream_t stream = cs;   // LAUNCH enqueues on `stream`
            cudaError_t le = launch_reset_state(ws.done, ws.iters, ws.success, ws.unsat, B, Bp, d->T, stream);
            d->prof.launches++;
            if (le != cudaSuccess) rc = fail(LDPC_ERR_CUDA, "reset (capture): %s", cudaGetErrorString(le));
            if (!rc) rc = run_span(d, ws, Bp, 0, d->T, want_post ? POST_EACH : POST_NONE, stream);
        }
        cudaGraph_t graph = nullptr;
        cudaError_t ee = cudaStreamEndCapture(cs, &graph);
        ldpc_decoder::GraphEntry ge;
        ge.launches = d->prof.launches - before.launches;
        ge.cn_launches = d->prof.cn_launches - before.cn_launches;
        ge.vn_launches = d->prof.vn_launches - before.vn_launches;
        d->prof = before;              // nothing ran yet: the replay below does the counting
        if (rc) {
            if (graph) cudaGraphDestroy(graph);
            return rc;
        }
        if (ee != cudaSuccess) return fail(LDPC_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(ee));
        ee = cudaGraphInstantiate(&ge.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ee != cudaSuccess) return fail(LDPC_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ee));
        ge.llrT = ws.llrT;
        ge.B = B;
        ge.Bp = Bp;
        ge.want_post = want_post;
        d->graphs.push_back(ge);
        hit = &d->graphs.back();
    }
    hit->stamp = ++d->graph_clock;
    CU(cudaGraphLaunch(hit->exec, stream));
    d->prof.launches += hit->launches;
    d->prof.cn_launches += hit->cn_launches;
    d->prof.vn_launches += hit->vn_launches;
    d->stat_graph_replays++;
    return LDPC_OK;
}

// The whole decode of the frames resident as llrT [n][Bp] in the context's root workspace, results delivered per
// OutSpec, as a RESUMABLE job: start() enqueues the first spans, resume() is called once the OLDEST outstanding
// checkpoint's count has arrived (wait_event()) and enqueues more.  The plain entry points drive one job to
// completion; the host pipeline alternates between two.
//
// Flooding with early stop runs in spans between checkpoints.  At a checkpoint the number of running frames
// comes back to the host: zero ends the decode (no empty launches up to T); if at most 60 % of the level's
// lanes still run, those frames' LLR and V2C columns are gathered into a dense child level that carries on
// from the same iteration, and the parent level delivers the results of its finished frames.  Every frame
// sees exactly the arithmetic of the uncompacted schedule (columns are independent), so results are identical.
//
// The host decides ONE SPAN LATE: while it waits for the count of checkpoint k, span k+1 (and its checkpoint) is
// already enqueued, so the GPU never idles for a host round trip.  A span that runs after every frame has
// stopped does nothing (all warps leave at once), so the late "all stopped" decision costs one empty span; a
// compaction first drains the one outstanding checkpoint, so the gather works on exact counts as before.
struct DecodeJob {
    ldpc_decoder* d = nullptr;
    ldpc_decoder::Ctx* cx = nullptr;
    cudaStream_t stream = nullptr;
    OutSpec o;
    bool want_post = false, checkpoints = false;
    int post_mode = POST_NONE;        // of the spans enqueued from now on
    Workspace* ws = nullptr;          // current level
    const int32_t* map = nullptr;     // its frames -> frames of the caller's batch
    int64_t curB = 0, curBp = 0;
    size_t level = 0;
    int t = 0, quiet = 0;
    uint32_t cp_issued = 0, cp_done = 0;   // checkpoints enqueued / consumed (slot = number & 1)
    int cp_t[2] = {0, 0};             // iteration count at each outstanding checkpoint
    int64_t last_pending = 0;         // running frames at the last consumed checkpoint (or at the start of the level)
    int last_t = 0;
    bool finished = false;            // everything up to the delivery of the results is enqueued
    bool waiting = false;             // the job needs resume() once wait_event() has completed
    cudaEvent_t wait_event() const { return cx->ev[cp_done & 1u]; }
};

// enqueue the next span; either the decode ends with it (results delivered, `finished`) or a checkpoint is recorded
int job_enqueue(DecodeJob& j) {
    ldpc_decoder* d = j.d;
    cudaStream_t stream = j.stream;
    const int t1 = j.checkpoints ? next_checkpoint(j.t, d->T, j.curBp, j.quiet, d->checkpoint_step) : d->T;
    int rc = run_span(d, *j.ws, j.curBp, j.t, t1, j.post_mode, stream);
    if (rc) return rc;
    j.t = t1;
    if (j.t >= d->T) {
        j.finished = true;
        return emit_level(d, *j.ws, j.curB, j.curBp, j.map, nullptr, j.o, stream);
    }
    const uint32_t slot = j.cp_issued & 1u;
    LAUNCH(K_OTHER, launch_pending_scan(j.ws->done, j.curBp, j.cx->d_scan, j.cx->d_total, j.cx->h_total_dev + slot, stream));
    CU(cudaEventRecord(j.cx->ev[slot], stream));
    j.cp_t[slot] = j.t;
    j.cp_issued++;
    return LDPC_OK;
}

// Keep one span in flight, or two.  A second span hides the host round trip of a checkpoint (the GPU idles for it: tens
// of microseconds per iteration at large batches), but a decision taken one span late runs that span at the old level
// size.  Measured (65 536 frames of the (16200,7200)-shaped code, tools/r02_probe.py / mc_knob_probe.py, r02ab): no frame
// ever stops, T = 10: 87.6 ms with one span, 88.7 with two; 3 dB, T = 50: 76.5 ms with one, 80.7 with two while nobody
// has stopped (LDPC_SPECULATE=-1), 88.7 with two throughout (=1).  The round trip is already down to a mapped-memory
// read behind an event, so one span stays the default.
int job_top_up(DecodeJob& j) {
    const bool nobody_stopped = j.level == 0 && j.last_pending >= j.curB;   // (child levels exist because frames stop)
    const uint32_t depth = (j.d->speculate > 0 || (j.d->speculate < 0 && nobody_stopped)) ? 2u : 1u;
    while (!j.finished && j.cp_issued - j.cp_done < depth) {
        int rc = job_enqueue(j);
        if (rc) return rc;
    }
    j.waiting = !j.finished;
    return LDPC_OK;
}

int job_start(DecodeJob& j, ldpc_decoder* d, ldpc_decoder::Ctx& cx, int64_t B, int64_t Bp, const OutSpec& o, cudaStream_t stream) {
    j = DecodeJob();
    j.d = d;
    j.cx = &cx;
    j.stream = stream;
    j.o = o;
    j.want_post = o.post != nullptr;
    Workspace& root = cx.root;
    if (j.want_post) {
        int rc = post_ensure(d, root);
        if (rc) return rc;
    }
    if (d->schedule == LDPC_SCHEDULE_LAYERED) {
        int rc = run_layered(d, root, B, Bp, j.want_post, stream);
        if (rc) return rc;
        return emit_level(d, root, B, Bp, nullptr, nullptr, o, stream);
    }
    {
        SmallLaunch sp{};
        if (fill_small(d, &root, B, Bp, j.want_post, sp)) {
            // small code: the whole decode in one launch, messages in shared memory
            LAUNCH(K_OTHER, launch_small_decode(d->dtype, sp, stream));
            d->stat_small++;
            return emit_level(d, root, B, Bp, nullptr, nullptr, o, stream);
        }
    }
    if (launch_bound(d, Bp)) {
        int rc = replay_graph(d, root, B, Bp, j.want_post, stream);
        if (rc) return rc;
        return emit_level(d, root, B, Bp, nullptr, nullptr, o, stream);
    }
    LAUNCH(K_OTHER, launch_reset_state(root.done, root.iters, root.success, root.unsat, B, Bp, d->T, stream));
    j.checkpoints = d->early_stop && d->compact && d->T > 2;
    j.ws = &root;
    j.curB = B;
    j.curBp = Bp;
    j.last_pending = B;
    if (j.want_post) {
        // with checkpoints the stop rate is known one span late and the mode follows it (job_resume); the first
        // iterations rarely stop a frame.  Without them: on-stop when nothing can stop early, else every iteration.
        if (d->post_mode) j.post_mode = d->post_mode;
        else j.post_mode = (j.checkpoints || !d->early_stop) ? POST_ON_STOP : POST_EACH;
    }
    if (j.checkpoints) {
        int rc = scan_ensure(cx, Bp);
        if (rc) return rc;
    }
    return job_top_up(j);
}

// precondition: j.waiting and wait_event() has completed
int job_resume(DecodeJob& j) {
    ldpc_decoder* d = j.d;
    ldpc_decoder::Ctx& cx = *j.cx;
    cudaStream_t stream = j.stream;
    const ldpc_graph* g = d->g;
    j.waiting = false;
    auto consume = [&]() -> int64_t {
        const uint32_t slot = j.cp_done & 1u;
        const int64_t pending = cx.h_total[slot];
        const int t_cp = j.cp_t[slot];
        j.cp_done++;
        j.quiet = (pending >= j.curB) ? j.quiet + 1 : 0;
        if (j.want_post && d->post_mode == 0 && t_cp > j.last_t && j.last_pending > 0) {
            // frames stopped per iteration, as a share of the running ones: the on-stop pass re-reads a whole lane /
            // sector for one newly stopped frame, so it only pays while stops are rare
            const double rate = (double)(j.last_pending - pending) / (double)j.last_pending / (double)(t_cp - j.last_t);
            j.post_mode = rate < 0.02 ? POST_ON_STOP : POST_EACH;
        }
        j.last_pending = pending;
        j.last_t = t_cp;
        return pending;
    };
    int64_t pending = consume();
    if (pending == 0) {
        d->stat_early_exits++;   // decisions (and posteriors) of stopped frames are already in place
        return emit_level(d, *j.ws, j.curB, j.curBp, j.map, nullptr, j.o, stream);
    }
    if (j.curBp >= d->compact_min_frames && pending * 100 <= j.curBp * d->compact_percent && j.level < kMaxLevels) {
        // ---- move the running frames to a dense child level ----
        if (j.cp_done != j.cp_issued) {   // the span in flight ends with a checkpoint of its own: gather on ITS count
            CU(cudaEventSynchronize(j.wait_event()));
            pending = consume();
            if (pending == 0) {
                d->stat_early_exits++;
                return emit_level(d, *j.ws, j.curB, j.curBp, j.map, nullptr, j.o, stream);
            }
        }
        if (cx.levels.size() <= j.level) cx.levels.emplace_back();   // capacity reserved at creation: no reallocation
        ldpc_decoder::Level& lv = cx.levels[j.level];
        const int64_t childBp = pad_frames(pending);
        const int64_t level_cap = pad_frames((j.curBp * d->compact_percent + 99) / 100);   // any later count of this level fits
        if (lv.cap < childBp) {
            const int64_t cap = std::max(childBp, level_cap);
            cudaFree(lv.idx);
            cudaFree(lv.map);
            lv.idx = lv.map = nullptr;
            lv.cap = 0;
            CU(cudaMalloc((void**)&lv.idx, (size_t)cap * sizeof(int32_t)));
            CU(cudaMalloc((void**)&lv.map, (size_t)cap * sizeof(int32_t)));
            lv.cap = cap;
        }
        int rc = ws_ensure(d, lv.ws, std::max(childBp, std::min(lv.cap, level_cap)));
        if (rc == LDPC_ERR_NOMEM) {   // no room for a child level: carry on uncompacted
            cudaGetLastError();
            lv.ws.release();
            return job_top_up(j);
        }
        if (rc) return rc;
        if (j.want_post) {
            rc = post_ensure(d, lv.ws);
            if (rc) return rc;
        }
        LAUNCH(K_OTHER, launch_pending_indices(j.ws->done, j.curBp, cx.d_scan, lv.idx, stream));
        LAUNCH(K_OTHER, launch_compose_map(lv.idx, j.map, lv.map, pending, stream));
        LAUNCH(K_OTHER, launch_gather_cols(d->dtype, j.ws->llrT, j.curBp, lv.ws.llrT, childBp, lv.idx, pending, g->n, stream));
        LAUNCH(K_OTHER, launch_gather_cols(d->dtype, j.ws->v2c, j.curBp, lv.ws.v2c, childBp, lv.idx, pending, g->E, stream));
        LAUNCH(K_OTHER, launch_reset_state(lv.ws.done, lv.ws.iters, lv.ws.success, lv.ws.unsat, pending, childBp, d->T, stream));
        // the parent delivers its stopped frames (their decisions / posteriors are in place)
        rc = emit_level(d, *j.ws, j.curB, j.curBp, j.map, j.ws->done, j.o, stream);
        if (rc) return rc;
        d->stat_compactions++;
        j.ws = &lv.ws;
        j.map = lv.map;
        j.curB = pending;
        j.curBp = childBp;
        ++j.level;
    }
    return job_top_up(j);
}

int job_drive(DecodeJob& j) {
    int rc = LDPC_OK;
    while (!rc && j.waiting) {
        CU(cudaEventSynchronize(j.wait_event()));
        rc = job_resume(j);
    }
    return rc;
}

int ctx_enter(ldpc_decoder::Ctx& cx, cudaStream_t stream) {
    if (!cx.tail) CU(cudaEventCreateWithFlags(&cx.tail, cudaEventDisableTiming));
    if (cx.tail_valid && cx.tail_stream != stream) CU(cudaStreamWaitEvent(stream, cx.tail, 0));
    return LDPC_OK;
}

int ctx_leave(ldpc_decoder::Ctx& cx, cudaStream_t stream) {
    CU(cudaEventRecord(cx.tail, stream));
    cx.tail_stream = stream;
    cx.tail_valid = true;
    return LDPC_OK;
}

int decode_resident(ldpc_decoder* d, ldpc_decoder::Ctx& cx, int64_t B, int64_t Bp, const OutSpec& o, cudaStream_t stream) {
    DecodeJob j;
    int rc = job_start(j, d, cx, B, Bp, o, stream);
    return rc ? rc : job_drive(j);
}

// pack + job_start on the context's root workspace
int job_start_on_device(DecodeJob& j, ldpc_decoder* d, ldpc_decoder::Ctx& cx, const void* llr, int64_t B, uint8_t* bits,
                        uint32_t* packed, void* post, int32_t* iters, uint8_t* success, cudaStream_t stream) {
    const int64_t Bp = pad_frames(B);
    {
        SmallLaunch sp{};
        if (fill_small(d, nullptr, B, Bp, post != nullptr, sp)) {
            // small code: the whole decode in one launch on the caller's row-major buffers (no workspace, no layout
            // conversion): LLR rows in, decisions / posteriors / iterations / success out
            sp.llr_rows = llr;
            sp.bits_rows = bits;
            sp.packed_rows = packed;
            sp.post_rows = post;
            sp.iters = iters;
            sp.success = success;
            d->prof.frames_padded = Bp;
            j = DecodeJob();
            LAUNCH(K_OTHER, launch_small_decode(d->dtype, sp, stream));
            d->stat_small++;
            return LDPC_OK;
        }
    }
    {
        ResidentLaunch rp{};
        if (fill_resident(d, B, rp)) {
            // messages of a frame fit one SM: one thread block per frame keeps them in shared memory for all T
            // iterations (row-major LLRs in, results out; no workspace, no message traffic through HBM)
            rp.llr_rows = static_cast<const float*>(llr);
            rp.bits_rows = bits;
            rp.packed_rows = packed;
            rp.post_rows = static_cast<float*>(post);
            rp.iters = iters;
            rp.success = success;
            d->prof.frames_padded = Bp;
            j = DecodeJob();
            LAUNCH(K_OTHER, launch_resident_decode(rp, stream));
            d->stat_resident++;
            return LDPC_OK;
        }
    }
    int rc = ws_ensure(d, cx.root, Bp);
    if (rc) return rc;
    d->prof.frames_padded = Bp;
    Workspace& ws = cx.root;
    LAUNCH(K_OTHER, launch_pack(d->dtype, llr, ws.llrT, B, Bp, d->g->n, ws.done, ws.iters, ws.success, d->T, stream));
    OutSpec o;
    o.bits = bits;
    o.packed = packed;
    o.post = post;
    o.iters = iters;
    o.success = success;
    return job_start(j, d, cx, B, Bp, o, stream);
}

// Chunk plan of the host pipeline: (offset, frames) pairs covering [0, B), each at most `chunk` frames.
// The link delivers a chunk's LLRs a little faster than the kernels consume them, so the first chunks grow
// geometrically (x1.5 from 1/8 of a chunk: the kernels start ~1 ms into the call); the rest of the batch is
// split into equal parts so that no small remainder is left for the un-overlapped tail.  Sizes are multiples of
// one CTA's frame block (256 lanes x V frames) where the batch allows it: a partly filled last block costs a
// whole CTA slot.
std::vector<std::pair<int64_t, int64_t>> plan_chunks(int64_t B, int64_t chunk, int V) {
    std::vector<std::pair<int64_t, int64_t>> chunks;
    if (chunk < 1 || chunk > B) chunk = B;
    const int64_t blk = (int64_t)256 * V;
    const int64_t align = (chunk % blk == 0 && B >= 4 * blk) ? blk : kFrameAlign;
    auto round_up = [&](int64_t x) { return (x + align - 1) / align * align; };
    int64_t off = 0;
    if (B > 2 * chunk) {
        for (int64_t c = std::max<int64_t>(align, chunk / 8); c < chunk && B - off > 2 * chunk; c = round_up(c * 3 / 2)) {
            chunks.emplace_back(off, c);
            off += c;
        }
    }
    while (off < B) {   // equal parts of at most `chunk` frames
        const int64_t left = B - off;
        const int64_t parts = (left + chunk - 1) / chunk;
        const int64_t b = std::min<int64_t>(std::min<int64_t>(round_up((left + parts - 1) / parts), chunk), left);
        chunks.emplace_back(off, b);
        off += b;
    }
    return chunks;
}

// With posteriors the outputs of a chunk are as large as its inputs, and nothing overlaps the copy-out of the LAST
// chunk: split it so that the un-overlapped tail is a quarter of it (halves, then quarters; multiples of 128 frames).
void taper_tail(std::vector<std::pair<int64_t, int64_t>>& chunks) {
    if (chunks.size() < 3) return;
    const std::pair<int64_t, int64_t> last = chunks.back();
    if (last.second < 8 * kFrameAlign) return;
    const int64_t half = last.second / 2 / kFrameAlign * kFrameAlign;
    const int64_t quarter = (last.second - half) / 2 / kFrameAlign * kFrameAlign;
    chunks.pop_back();
    chunks.emplace_back(last.first, half);
    chunks.emplace_back(last.first + half, quarter);
    chunks.emplace_back(last.first + half + quarter, last.second - half - quarter);
}

int decode_on_device(ldpc_decoder* d, ldpc_decoder::Ctx& cx, const void* llr, int64_t B, uint8_t* bits, uint32_t* packed,
                     void* post, int32_t* iters, uint8_t* success, cudaStream_t stream) {
    DecodeJob j;
    int rc = job_start_on_device(j, d, cx, llr, B, bits, packed, post, iters, success, stream);
    return rc ? rc : job_drive(j);
}

}  // namespace

extern "C" int ldpc_decoder_create(ldpc_graph* g, const ldpc_decoder_config* cfg, ldpc_decoder** out) {
    if (!out) return fail(LDPC_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!g || !cfg) return fail(LDPC_ERR_INVALID, "NULL argument");
    if (cfg->struct_size != (int32_t)sizeof(ldpc_decoder_config))
        return fail(LDPC_ERR_INVALID, "config struct_size %d != %d", cfg->struct_size, (int)sizeof(ldpc_decoder_config));
    if (cfg->dtype != LDPC_F32 && cfg->dtype != LDPC_F64) return fail(LDPC_ERR_INVALID, "bad dtype");
    if (cfg->max_iterations < 1) return fail(LDPC_ERR_INVALID, "max_iterations must be >= 1");
    if (cfg->n_beta < 0 || cfg->n_alpha < 0) return fail(LDPC_ERR_INVALID, "negative table width");
    if (cfg->n_beta > 0 && !cfg->beta) return fail(LDPC_ERR_INVALID, "beta is NULL");
    if (cfg->n_alpha > 0 && !cfg->alpha) return fail(LDPC_ERR_INVALID, "alpha is NULL");
    if (cfg->bc != 0) {
        if (cfg->bc < 2 || cfg->bc > 8) return fail(LDPC_ERR_UNSUPPORTED, "bc must be 0 or 2..8");
        if (cfg->dtype != LDPC_F32) return fail(LDPC_ERR_UNSUPPORTED, "quantised decoding is float32 only");
        if (cfg->n_quantizers < 1 || !cfg->thresholds || !cfg->quantizer_of_iter)
            return fail(LDPC_ERR_INVALID, "quantiser tables missing");
        if (((int64_t)cfg->n_quantizers << cfg->bc) > kMaxLutFloats)
            return fail(LDPC_ERR_UNSUPPORTED, "too many quantiser levels");
    }
    if (cfg->check_rule != LDPC_RULE_NORMALIZED && cfg->check_rule != LDPC_RULE_OFFSET)
        return fail(LDPC_ERR_INVALID, "bad check_rule");
    if (cfg->check_rule == LDPC_RULE_OFFSET && cfg->bc != 0)
        return fail(LDPC_ERR_UNSUPPORTED, "the offset rule has no quantised variant in the reference");
    if (cfg->schedule != LDPC_SCHEDULE_FLOODING && cfg->schedule != LDPC_SCHEDULE_LAYERED)
        return fail(LDPC_ERR_INVALID, "bad schedule");
    if (cfg->schedule == LDPC_SCHEDULE_LAYERED) {
        if (cfg->bc == 0 || cfg->n_beta || cfg->n_alpha || cfg->check_rule != LDPC_RULE_NORMALIZED)
            return fail(LDPC_ERR_UNSUPPORTED, "the layered schedule exists for plain RCQ only (rcq_decoder.py:281-350)");
        if (g->nonempty_checks == 1)
            return fail(LDPC_ERR_UNSUPPORTED, "layered schedule on a graph with a single non-empty check");
    }
    if (cfg->dtype == LDPC_F64 && g->max_dv > 129)
        return fail(LDPC_ERR_UNSUPPORTED, "float64 path models np.sum up to 128 terms (max variable degree 129)");
    const int T = cfg->max_iterations;
    const int64_t E = g->E;
    if (cfg->beta_index)
        for (int64_t e = 0; e < E; ++e)
            if (cfg->beta_index[e] < 0 || cfg->beta_index[e] >= std::max(cfg->n_beta, 1))
                return fail(LDPC_ERR_INVALID, "beta_index[%lld] out of range", (long long)e);
    if (cfg->alpha_index)
        for (int32_t j = 0; j < g->n; ++j)
            if (cfg->alpha_index[j] < 0 || cfg->alpha_index[j] >= std::max(cfg->n_alpha, 1))
                return fail(LDPC_ERR_INVALID, "alpha_index[%d] out of range", j);

    ldpc_decoder* d = new (std::nothrow) ldpc_decoder();
    if (!d) return fail(LDPC_ERR_NOMEM, "host allocation failed");
    d->g = g;
    d->dtype = cfg->dtype;
    d->V = cfg->dtype == LDPC_F32 ? 4 : 2;
    d->rsz = cfg->dtype == LDPC_F32 ? 4 : 8;
    d->T = T;
    d->early_stop = cfg->early_stop ? 1 : 0;
    d->n_beta = cfg->n_beta;
    d->n_alpha = cfg->n_alpha;
    d->bc = cfg->bc;
    d->Q = cfg->bc ? cfg->n_quantizers : 0;
    d->nth = cfg->bc ? (1 << (cfg->bc - 1)) : 0;
    d->check_rule = cfg->check_rule;
    d->schedule = cfg->schedule;
    for (auto& cx : d->cx) cx.levels.reserve(kMaxLevels);   // jobs keep pointers into these vectors
    if (const char* hc = getenv("LDPC_HOST_CHUNK")) d->host_chunk = atoll(hc);
    if (const char* hd = getenv("LDPC_HOST_DUAL")) d->host_dual = atoi(hd) != 0;
    if (const char* lp = getenv("LDPC_LAYERED_PIPE")) d->layered_pipe = atoi(lp) != 0;
    if (const char* lv = getenv("LDPC_LAYERED_V2_FRAMES")) d->layered_v2_frames = atoll(lv);
    if (const char* lv = getenv("LDPC_LAYERED_V4_FRAMES")) d->layered_v4_frames = atoll(lv);
    if (const char* ll = getenv("LDPC_LAYERED_LEVELS")) d->layered_levels = atoi(ll) != 0;
    if (const char* ls = getenv("LDPC_LAYERED_STAGE")) d->layered_stage = std::min(2, std::max(0, atoi(ls)));  // tuning knob: frames per pipeline chunk
    if (const char* wr = getenv("LDPC_WIDE_RING")) d->wide_ring = atoi(wr) != 0; // A/B switch for the wide-check kernel
    if (const char* cp = getenv("LDPC_COMPACT")) d->compact = atoi(cp) != 0;      // A/B switch for frame compaction
    if (const char* fi = getenv("LDPC_FINE_ITEMS_MAX_FRAMES")) d->fine_items_max_frames = atoll(fi);
    if (const char* gr = getenv("LDPC_GRAPHS")) d->use_graphs = atoi(gr) != 0;
    if (const char* cf = getenv("LDPC_COMPACT_PERCENT")) d->compact_percent = std::min(95, std::max(5, atoi(cf)));
    if (const char* cs = getenv("LDPC_CHECKPOINT_STEP")) d->checkpoint_step = std::max(1, atoi(cs));
    if (const char* sp = getenv("LDPC_SPECULATE")) d->speculate = atoi(sp);
    if (const char* tg = getenv("LDPC_TRAIN_GENERAL")) d->train_general = atoi(tg) != 0;
    if (const char* sm = getenv("LDPC_SMALL")) d->use_small = atoi(sm) != 0;
    if (const char* rs = getenv("LDPC_RESIDENT")) d->use_resident = atoi(rs) != 0;
    if (const char* pm = getenv("LDPC_POST_MODE")) d->post_mode = std::min(2, std::max(0, atoi(pm)));
    if (const char* cm = getenv("LDPC_COMPACT_MIN_FRAMES")) d->compact_min_frames = std::max<int64_t>(atoll(cm), kFrameAlign);

    DeviceGuard guard(g->device);
    if (!guard.ok) {
        delete d;
        return fail(LDPC_ERR_CUDA, "cannot select CUDA device %d", g->device);
    }
    int rc = LDPC_OK;
    cudaDeviceGetAttribute(&d->sm_count, cudaDevAttrMultiProcessorCount, g->device);
    if (const char* rm = getenv("LDPC_RESIDENT_MAX_FRAMES")) d->resident_max_frames = atoll(rm);
    if (cfg->n_beta > 0 && cfg->beta_index) {
        std::vector<int32_t> bidx((size_t)E);
        for (int64_t e = 0; e < E; ++e) bidx[(size_t)g->slot_of_edge[(size_t)e]] = cfg->beta_index[e];
        for (const WorkItem& it : g->cn[0].items)   // does any check mix columns?
            for (int c = 0; c < it.count && !d->beta_per_edge; ++c)
                for (int k = 1; k < it.deg; ++k)
                    if (bidx[(size_t)it.first_slot + (size_t)c * it.deg + k] != bidx[(size_t)it.first_slot + (size_t)c * it.deg]) {
                        d->beta_per_edge = 1;
                        break;
                    }
        rc = upload(&d->d_bidx, bidx);
        if (!rc && g->res.ok) {   // the same columns in the slot order of the CTA-resident decode
            std::vector<int32_t> rb((size_t)g->res.E_phys, 0);
            for (int64_t sl = 0; sl < E; ++sl) rb[(size_t)g->res.phys[(size_t)sl]] = bidx[(size_t)sl];
            rc = upload(&d->d_res_bidx, rb);
        }
    }
    if (!rc && cfg->n_alpha > 0 && cfg->alpha_index) {
        std::vector<int32_t> aidx((size_t)g->n);
        for (int32_t p = 0; p < g->n; ++p) aidx[(size_t)p] = cfg->alpha_index[g->vpos_var[(size_t)p]];
        rc = upload(&d->d_aidx, aidx);
        if (!rc && cfg->check_rule == LDPC_RULE_OFFSET) {
            std::vector<int32_t> as((size_t)E);
            for (int64_t sl = 0; sl < E; ++sl) as[(size_t)sl] = cfg->alpha_index[g->slot_var[(size_t)sl]];
            rc = upload(&d->d_aidx_slot, as);
            if (!rc && g->res.ok) {
                std::vector<int32_t> ra((size_t)g->res.E_phys, 0);
                for (int64_t sl = 0; sl < E; ++sl) ra[(size_t)g->res.phys[(size_t)sl]] = as[(size_t)sl];
                rc = upload(&d->d_res_aidx_slot, ra);
            }
        }
    }
    if (!rc && cfg->n_beta > 0) {
        size_t bytes = (size_t)T * cfg->n_beta * d->rsz;
        cudaError_t e = cudaMalloc(&d->d_beta, bytes);
        if (e == cudaSuccess) e = cudaMemcpy(d->d_beta, cfg->beta, bytes, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) rc = fail(LDPC_ERR_CUDA, "beta upload: %s", cudaGetErrorString(e));
    }
    if (!rc && cfg->n_alpha > 0) {
        size_t bytes = (size_t)T * cfg->n_alpha * d->rsz;
        cudaError_t e = cudaMalloc(&d->d_alpha, bytes);
        if (e == cudaSuccess) e = cudaMemcpy(d->d_alpha, cfg->alpha, bytes, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) rc = fail(LDPC_ERR_CUDA, "alpha upload: %s", cudaGetErrorString(e));
    }
    if (!rc && cfg->bc) {
        const int nth = d->nth, nl = 1 << cfg->bc;
        std::vector<float> thr(cfg->thresholds, cfg->thresholds + (size_t)d->Q * nth);
        std::vector<float> lut((size_t)d->Q * nl);
        d->mono.assign(d->Q, 1);
        for (int q = 0; q < d->Q; ++q) {
            for (int j = 1; j < nth; ++j)
                if (!(thr[(size_t)q * nth + j] >= thr[(size_t)q * nth + j - 1])) d->mono[q] = 0;
            for (int code = 0; code < nl; ++code) {
                // rcq_decoder.py:106-119: sign = 1 - 2*sign_bit, value = sign * threshold[idx]
                int sb = code >= nth;
                float sign = 1.f - 2.f * (float)sb;
                lut[(size_t)q * nl + code] = sign * thr[(size_t)q * nth + (code % nth)];
            }
        }
        d->q_of_iter.assign(cfg->quantizer_of_iter, cfg->quantizer_of_iter + T);
        for (int t = 0; t < T && !rc; ++t)
            if (d->q_of_iter[t] < 0 || d->q_of_iter[t] >= d->Q) rc = fail(LDPC_ERR_INVALID, "quantizer_of_iter[%d] out of range", t);
        if (!rc) rc = upload(&d->d_thr, thr);
        if (!rc) rc = upload(&d->d_lut, lut);
        if (!rc) rc = upload(&d->d_qoi, d->q_of_iter);
        if (!rc) {
            std::vector<int32_t> mono(d->mono.begin(), d->mono.end());
            for (int mq : d->mono) d->all_mono &= mq;
            rc = upload(&d->d_mono, mono);
        }
    }
    if (rc) {
        ldpc_decoder_destroy(d);
        return rc;
    }
    *out = d;
    return LDPC_OK;
}

extern "C" int ldpc_decoder_set_weights(ldpc_decoder* d, const void* beta, const void* alpha) {
    if (!d) return fail(LDPC_ERR_INVALID, "NULL decoder");
    DeviceGuard guard(d->g->device);
    // decodes return with kernels still in flight on the caller's stream: the tables must not change under them
    for (auto& cx : d->cx)
        if (cx.tail_valid) CU(cudaEventSynchronize(cx.tail));
    if (beta) {
        if (!d->d_beta) return fail(LDPC_ERR_INVALID, "decoder was created without beta");
        CU(cudaMemcpy(d->d_beta, beta, (size_t)d->T * d->n_beta * d->rsz, cudaMemcpyHostToDevice));
    }
    if (alpha) {
        if (!d->d_alpha) return fail(LDPC_ERR_INVALID, "decoder was created without alpha");
        CU(cudaMemcpy(d->d_alpha, alpha, (size_t)d->T * d->n_alpha * d->rsz, cudaMemcpyHostToDevice));
    }
    return LDPC_OK;
}

extern "C" int ldpc_decoder_destroy(ldpc_decoder* d) {
    if (!d) return LDPC_OK;
    DeviceGuard guard(d->g->device);
    cudaDeviceSynchronize();
    d->pipe.release();
    for (auto& ge : d->graphs)
        if (ge.exec) cudaGraphExecDestroy(ge.exec);
    if (d->cap_stream) cudaStreamDestroy(d->cap_stream);
    for (auto& cx : d->cx) {
        cx.root.release();
        for (auto& lv : cx.levels) {
            lv.ws.release();
            cudaFree(lv.idx);
            cudaFree(lv.map);
        }
        cudaFree(cx.d_scan);
        cudaFree(cx.d_total);
        if (cx.h_total) cudaFreeHost(cx.h_total);
        for (auto& e : cx.ev)
            if (e) cudaEventDestroy(e);
        if (cx.tail) cudaEventDestroy(cx.tail);
    }
    for (auto& ev : d->ev_pool) {
        cudaEventDestroy(ev.first);
        cudaEventDestroy(ev.second);
    }
    d->train.ws.release();
    cudaFree(d->train.v2c_hist);
    cudaFree(d->train.c2v_hist);
    cudaFree(d->train.g_v2c);
    cudaFree(d->train.g_c2v);
    cudaFree(d->train.g_post);
    cudaFree(d->train.g_part);
    cudaFree(d->d_bidx);
    cudaFree(d->d_aidx);
    cudaFree(d->d_aidx_slot);
    cudaFree(d->d_res_bidx);
    cudaFree(d->d_res_aidx_slot);
    cudaFree(d->d_beta);
    cudaFree(d->d_alpha);
    cudaFree(d->d_thr);
    cudaFree(d->d_lut);
    cudaFree(d->d_qoi);
    cudaFree(d->d_mono);
    delete d;
    return LDPC_OK;
}

extern "C" int ldpc_decoder_reserve(ldpc_decoder* d, int64_t frames) {
    if (!d || frames < 1) return fail(LDPC_ERR_INVALID, "bad arguments");
    DeviceGuard guard(d->g->device);
    return ws_ensure(d, d->cx[0].root, pad_frames(frames));
}

namespace {
int decode_device_impl(ldpc_decoder* d, const void* llr, int64_t B, uint8_t* bits, uint32_t* packed, void* posterior,
                       int32_t* iterations, uint8_t* success, void* stream) {
    if (!d || !llr) return fail(LDPC_ERR_INVALID, "NULL argument");
    if (B < 1) return fail(LDPC_ERR_INVALID, "B must be >= 1");
    DeviceGuard guard(d->g->device);
    if (!guard.ok) return fail(LDPC_ERR_CUDA, "cannot select CUDA device %d", d->g->device);
    int rc = ctx_enter(d->cx[0], (cudaStream_t)stream);
    if (!rc) rc = decode_on_device(d, d->cx[0], llr, B, bits, packed, posterior, iterations, success, (cudaStream_t)stream);
    if (!rc) rc = ctx_leave(d->cx[0], (cudaStream_t)stream);
    return rc;
}
int decode_host_impl(ldpc_decoder* d, const void* llr, int64_t B, uint8_t* bits, uint32_t* packed, void* posterior,
                     int32_t* iterations, uint8_t* success);
}  // namespace

extern "C" int ldpc_decode_device(ldpc_decoder* d, const void* llr, int64_t B, uint8_t* bits, void* posterior,
                                  int32_t* iterations, uint8_t* success, void* stream) {
    return decode_device_impl(d, llr, B, bits, nullptr, posterior, iterations, success, stream);
}

extern "C" int ldpc_decode_device_packed(ldpc_decoder* d, const void* llr, int64_t B, uint32_t* bits_packed, void* posterior,
                                         int32_t* iterations, uint8_t* success, void* stream) {
    return decode_device_impl(d, llr, B, nullptr, bits_packed, posterior, iterations, success, stream);
}

extern "C" int ldpc_decode_host(ldpc_decoder* d, const void* llr, int64_t B, uint8_t* bits, void* posterior,
                                int32_t* iterations, uint8_t* success) {
    return decode_host_impl(d, llr, B, bits, nullptr, posterior, iterations, success);
}

extern "C" int ldpc_decode_host_packed(ldpc_decoder* d, const void* llr, int64_t B, uint32_t* bits_packed, void* posterior,
                                       int32_t* iterations, uint8_t* success) {
    return decode_host_impl(d, llr, B, nullptr, bits_packed, posterior, iterations, success);
}

// Host-buffer entry points: chunked pipeline with two decode jobs in flight (see HostPipe, DecodeJob).  The decisions
// travel back either as one byte per bit (`bits`) or as packed rows (`packed`, an eighth of the bytes).
namespace {
int decode_host_impl(ldpc_decoder* d, const void* llr, int64_t B, uint8_t* bits, uint32_t* packed, void* posterior,
                     int32_t* iterations, uint8_t* success) {
    if (!d || !llr) return fail(LDPC_ERR_INVALID, "NULL argument");
    if (B < 1) return fail(LDPC_ERR_INVALID, "B must be >= 1");
    DeviceGuard guard(d->g->device);
    if (!guard.ok) return fail(LDPC_ERR_CUDA, "cannot select CUDA device %d", d->g->device);
    if (bits && packed) return fail(LDPC_ERR_INVALID, "one decision format per call");
    const ldpc_graph* g = d->g;
    const int64_t n = g->n;
    const int64_t row_words = (n + 31) / 32;
    HostPipe& pp = d->pipe;
    // Two chunks decode at a time, so 2 x 4096 frames fill the GPU, and the smaller the chunk the shorter the ramp and
    // the tail (65 536 frames of the (16200,7200)-shaped code, tools/e2e_trace.py: 98.3 ms with 8192-frame chunks, 94.2
    // with 4096 or 2048).  With posteriors the copy-out is as large as the copy-in and the link, busy in both directions,
    // is what bounds the call: small chunks get the output flowing early and keep it flowing (116.2 ms with 8192-frame
    // chunks, 105.3 with 4096, 103.3 with 2048, 107.0 with 1024).
    int64_t chunk = d->host_chunk > 0 ? d->host_chunk : (posterior ? 2048 : 4096);
    if (B <= chunk) chunk = B;
    if (!pp.s_in) {
        CU(cudaStreamCreateWithFlags(&pp.s_in, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&pp.s_run, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&pp.s_run2, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&pp.s_out, cudaStreamNonBlocking));
        for (auto& b : pp.buf) {
            CU(cudaEventCreateWithFlags(&b.in_ready, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&b.run_done, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&b.out_done, cudaEventDisableTiming));
        }
    }
    if (pp.cap < chunk || (posterior && !pp.post_cap)) {
        for (auto& b : pp.buf) {
            cudaFree(b.d_llr); cudaFree(b.d_bits); cudaFree(b.d_post); cudaFree(b.d_it); cudaFree(b.d_su);
            b.d_llr = nullptr; b.d_bits = nullptr; b.d_post = nullptr; b.d_it = nullptr; b.d_su = nullptr;
            CU(cudaMalloc(&b.d_llr, (size_t)chunk * n * d->rsz));
            CU(cudaMalloc((void**)&b.d_bits, (size_t)chunk * std::max<int64_t>(n, 4 * row_words)));   // either decision format
            if (posterior) CU(cudaMalloc(&b.d_post, (size_t)chunk * n * d->rsz));
            CU(cudaMalloc((void**)&b.d_it, (size_t)chunk * sizeof(int32_t)));
            CU(cudaMalloc((void**)&b.d_su, (size_t)chunk));
        }
        pp.cap = chunk;
        pp.post_cap = posterior != nullptr;
    }
    int rc = LDPC_OK;
    std::vector<std::pair<int64_t, int64_t>> chunks = plan_chunks(B, chunk, d->V);   // (offset, frames)
    if (posterior) taper_tail(chunks);
    // Two chunks decode at a time (two contexts, two kernel streams): each is a resumable job, the host enqueues
    // one span per job in turn and blocks on the older of the two outstanding checkpoints, so that while one job
    // waits for its host round trip -- and while its kernels drain -- the other one keeps the SMs busy.  The
    // input copy of chunk i+1 is enqueued before chunk i starts, so the copy engine never waits for the host.
    const int njobs = (d->host_dual && chunks.size() > 2) ? 2 : 1;
    cudaStream_t run_stream[2] = {pp.s_run, pp.s_run2};
    struct Slot {
        DecodeJob job;
        size_t chunk = 0;
        bool busy = false;
        uint64_t cp_seq = 0;   // order of the outstanding checkpoint among both jobs
    } slot[2];
    uint64_t seq = 0;
    const bool trace = getenv("LDPC_PIPE_TRACE") != nullptr;   // debug: per-chunk timeline on stderr
    std::vector<cudaEvent_t> tin(trace ? chunks.size() : 0, nullptr);
    auto enqueue_input = [&](size_t i) -> int {
        HostPipe::Buf& bf = pp.buf[i % kPipeDepth];
        if (i >= (size_t)kPipeDepth) CU(cudaStreamWaitEvent(pp.s_in, bf.run_done, 0));     // staging input consumed
        CU(cudaMemcpyAsync(bf.d_llr, (const char*)llr + (size_t)chunks[i].first * n * d->rsz,
                           (size_t)chunks[i].second * n * d->rsz, cudaMemcpyHostToDevice, pp.s_in));
        CU(cudaEventRecord(bf.in_ready, pp.s_in));
        if (trace) {
            cudaEventCreate(&tin[i]);
            cudaEventRecord(tin[i], pp.s_in);
        }
        return LDPC_OK;
    };
    std::vector<cudaEvent_t> tev(trace ? 1 + 3 * chunks.size() : 0, nullptr);
    auto mark = [&](size_t idx, cudaStream_t st) {
        if (!trace) return;
        cudaEventCreate(&tev[idx]);
        cudaEventRecord(tev[idx], st);
    };
    // everything of this slot's chunk is enqueued: hand its outputs to the copy-out stream
    auto finish_chunk = [&](int sidx) -> int {
        Slot& sl = slot[sidx];
        const size_t i = sl.chunk;
        const int64_t off = chunks[i].first, b = chunks[i].second;
        HostPipe::Buf& bf = pp.buf[i % kPipeDepth];
        cudaStream_t rs = run_stream[sidx];
        CU(cudaEventRecord(bf.run_done, rs));
        mark(2 + 3 * i, rs);
        CU(cudaStreamWaitEvent(pp.s_out, bf.run_done, 0));
        if (bits) CU(cudaMemcpyAsync(bits + (size_t)off * n, bf.d_bits, (size_t)b * n, cudaMemcpyDeviceToHost, pp.s_out));
        if (packed)
            CU(cudaMemcpyAsync(packed + (size_t)off * row_words, bf.d_bits, (size_t)b * row_words * sizeof(uint32_t),
                               cudaMemcpyDeviceToHost, pp.s_out));
        if (posterior)
            CU(cudaMemcpyAsync((char*)posterior + (size_t)off * n * d->rsz, bf.d_post, (size_t)b * n * d->rsz,
                               cudaMemcpyDeviceToHost, pp.s_out));
        if (iterations) CU(cudaMemcpyAsync(iterations + off, bf.d_it, (size_t)b * sizeof(int32_t), cudaMemcpyDeviceToHost, pp.s_out));
        if (success) CU(cudaMemcpyAsync(success + off, bf.d_su, (size_t)b, cudaMemcpyDeviceToHost, pp.s_out));
        CU(cudaEventRecord(bf.out_done, pp.s_out));
        mark(3 + 3 * i, pp.s_out);
        sl.busy = false;
        return LDPC_OK;
    };
    // resume the job with the OLDER outstanding checkpoint until slot `target` is free (target < 0: all slots)
    auto pump = [&](int target) -> int {
        while (true) {
            if (target >= 0 ? !slot[target].busy : (!slot[0].busy && !slot[1].busy)) return LDPC_OK;
            int pick = -1;
            for (int k = 0; k < 2; ++k)
                if (slot[k].busy && (pick < 0 || slot[k].cp_seq < slot[pick].cp_seq)) pick = k;
            Slot& sl = slot[pick];
            CU(cudaEventSynchronize(sl.job.wait_event()));
            int r = job_resume(sl.job);
            if (r) return r;
            if (sl.job.waiting) sl.cp_seq = ++seq;
            else if ((r = finish_chunk(pick))) return r;
        }
    };
    mark(0, pp.s_in);
    rc = enqueue_input(0);
    for (size_t i = 0; i < chunks.size() && !rc; ++i) {
        const int sidx = (int)(i % (size_t)njobs);
        rc = pump(sidx);   // the chunk that used this slot is through
        if (rc) break;
        if (i + 1 < chunks.size()) {
            // buffer (i+1) % depth was last used by chunk i+1-depth, finished at least one pump ago
            rc = enqueue_input(i + 1);
            if (rc) break;
        }
        HostPipe::Buf& bf = pp.buf[i % kPipeDepth];
        cudaStream_t rs = run_stream[sidx];
        CU(cudaStreamWaitEvent(rs, bf.in_ready, 0));
        if (i >= (size_t)kPipeDepth) CU(cudaStreamWaitEvent(rs, bf.out_done, 0));    // staging outputs copied out
        mark(1 + 3 * i, rs);
        Slot& sl = slot[sidx];
        sl.chunk = i;
        sl.busy = true;
        rc = job_start_on_device(sl.job, d, d->cx[1 + sidx], bf.d_llr, chunks[i].second, bits ? bf.d_bits : nullptr,
                                 packed ? reinterpret_cast<uint32_t*>(bf.d_bits) : nullptr,
                                 posterior ? bf.d_post : nullptr, bf.d_it, bf.d_su, rs);
        if (rc) break;
        if (sl.job.waiting) sl.cp_seq = ++seq;
        else rc = finish_chunk(sidx);
    }
    if (!rc) rc = pump(-1);
    for (cudaStream_t st : {pp.s_in, pp.s_run, pp.s_run2, pp.s_out}) {
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess && !rc) rc = fail(LDPC_ERR_CUDA, "stream sync: %s", cudaGetErrorString(e));
    }
    if (trace && !rc) {
        for (size_t i = 0; i < chunks.size(); ++i) {
            float a = 0, b = 0, c = 0;
            cudaEventElapsedTime(&a, tev[0], tev[1 + 3 * i]);
            cudaEventElapsedTime(&b, tev[0], tev[2 + 3 * i]);
            cudaEventElapsedTime(&c, tev[0], tev[3 + 3 * i]);
            float in = 0;
            cudaEventElapsedTime(&in, tev[0], tin[i]);
            fprintf(stderr, "chunk %2zu frames %6lld: input landed %7.2f ms  decode start %7.2f ms  end %7.2f ms (%.2f)  outputs copied %7.2f ms\n",
                    i, (long long)chunks[i].second, in, a, b, b - a, c);
        }
    }
    for (cudaEvent_t e : tev)
        if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : tin)
        if (e) cudaEventDestroy(e);
    return rc;
}
}  // namespace

extern "C" int ldpc_host_chunk_plan(int64_t B, int64_t chunk, int32_t frames_per_lane, int64_t* frames_out, int32_t max_chunks,
                                    int32_t* n_chunks) {
    if (B < 1 || !n_chunks || frames_per_lane < 1) return fail(LDPC_ERR_INVALID, "bad arguments");
    const auto plan = plan_chunks(B, chunk > 0 ? chunk : 4096, frames_per_lane);
    *n_chunks = (int32_t)plan.size();
    for (int32_t i = 0; i < (int32_t)plan.size() && i < max_chunks && frames_out; ++i) frames_out[i] = plan[(size_t)i].second;
    return LDPC_OK;
}

// =================================================================================================
// Posterior training
// =================================================================================================
namespace {
int train_supported(const ldpc_decoder* d) {
    if (d->dtype != LDPC_F32 || d->bc != 0 || d->check_rule != LDPC_RULE_NORMALIZED || d->schedule != LDPC_SCHEDULE_FLOODING)
        return fail(LDPC_ERR_UNSUPPORTED, "training exists for the float32 normalised neural min-sum decoders (flooding)");
    return LDPC_OK;
}
}  // namespace

extern "C" int ldpc_train_forward(ldpc_decoder* d, const float* llr, int64_t B, uint8_t* bits, float* posterior,
                                  int32_t* iterations, uint8_t* success, void* stream_) {
    if (!d || !llr) return fail(LDPC_ERR_INVALID, "NULL argument");
    if (B < 1) return fail(LDPC_ERR_INVALID, "B must be >= 1");
    int rc = train_supported(d);
    if (rc) return rc;
    DeviceGuard guard(d->g->device);
    if (!guard.ok) return fail(LDPC_ERR_CUDA, "cannot select CUDA device %d", d->g->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const ldpc_graph* g = d->g;
    ldpc_decoder::TrainCtx& tc = d->train;
    const int64_t Bp = pad_frames(B);
    const size_t E = (size_t)std::max<int64_t>(g->E, 1);
    tc.B = tc.Bp = 0;
    if (tc.cap < Bp) {
        cudaFree(tc.v2c_hist); cudaFree(tc.c2v_hist); cudaFree(tc.g_v2c); cudaFree(tc.g_c2v); cudaFree(tc.g_post);
        tc.v2c_hist = tc.c2v_hist = tc.g_v2c = tc.g_c2v = tc.g_post = nullptr;
        tc.cap = 0;
        rc = ws_ensure(d, tc.ws, Bp);
        if (rc) return rc;
        CU(cudaMalloc((void**)&tc.v2c_hist, (size_t)d->T * E * Bp * sizeof(float)));
        CU(cudaMalloc((void**)&tc.c2v_hist, (size_t)d->T * E * Bp * sizeof(float)));
        CU(cudaMalloc((void**)&tc.g_v2c, E * Bp * sizeof(float)));
        CU(cudaMalloc((void**)&tc.g_c2v, E * Bp * sizeof(float)));
        CU(cudaMalloc((void**)&tc.g_post, (size_t)g->n * Bp * sizeof(float)));
        tc.cap = tc.ws.cap;
    }
    Workspace& ws = tc.ws;
    const int64_t cap = ws.cap;   // row stride of the history slices is the launch's Bp, slices are sized for cap
    (void)cap;
    rc = post_ensure(d, ws);
    if (rc) return rc;
    d->prof.frames_padded = Bp;
    LAUNCH(K_OTHER, launch_pack(d->dtype, llr, ws.llrT, B, Bp, g->n, ws.done, ws.iters, ws.success, d->T, stream));
    LAUNCH(K_OTHER, launch_reset_state(ws.done, ws.iters, ws.success, ws.unsat, B, Bp, d->T, stream));
    const int64_t Wn = Bp / 32;
    const size_t slice = E * (size_t)Bp;
    for (int t = 0; t < d->T; ++t) {
        CnLaunch cn{};
        fill_cn(d, ws, Bp, t, cn);
        cn.src = (t == 0) ? ws.llrT : (const void*)(tc.v2c_hist + (size_t)t * slice);
        cn.dst = tc.c2v_hist + (size_t)t * slice;
        LAUNCH(K_CN, launch_cn(d->dtype, cn, stream));
        const bool last = (t == d->T - 1);
        VnLaunch vn{};
        fill_vn(d, ws, Bp, t, last, true, vn);
        vn.c2v = tc.c2v_hist + (size_t)t * slice;
        vn.v2c = last ? nullptr : (void*)(tc.v2c_hist + (size_t)(t + 1) * slice);
        LAUNCH(K_VN, launch_vn(d->dtype, vn, stream));
        uint32_t* cur = ws.unsat + (size_t)(t & 1) * Wn;
        uint32_t* nxt = ws.unsat + (size_t)((t + 1) & 1) * Wn;
        SynLaunch sy{};
        sy.hardw = ws.hardw;
        sy.Wn = Wn;
        sy.slot_var = g->d_slot_var;
        sy.items = g->cn[item_set(d, Bp)].d;
        sy.n_items = (int)g->cn[item_set(d, Bp)].items.size();
        sy.unsat = cur;
        LAUNCH(K_OTHER, launch_syndrome(sy, stream));
        LAUNCH(K_OTHER, launch_commit(d->V, cur, nxt, ws.done, ws.iters, ws.success, t + 1, Bp, stream));
    }
    OutSpec o;
    o.bits = bits;
    o.post = posterior;
    o.iters = iterations;
    o.success = success;
    rc = emit_level(d, ws, B, Bp, nullptr, nullptr, o, stream);
    if (rc) return rc;
    tc.B = B;
    tc.Bp = Bp;
    return LDPC_OK;
}

extern "C" int ldpc_train_backward(ldpc_decoder* d, const float* grad_posterior, float* grad_beta, float* grad_alpha,
                                   void* stream_) {
    if (!d || !grad_posterior) return fail(LDPC_ERR_INVALID, "NULL argument");
    int rc = train_supported(d);
    if (rc) return rc;
    ldpc_decoder::TrainCtx& tc = d->train;
    if (tc.B < 1) return fail(LDPC_ERR_INVALID, "ldpc_train_backward needs a preceding ldpc_train_forward");
    if (grad_beta && !d->d_beta) return fail(LDPC_ERR_INVALID, "decoder was created without beta");
    if (grad_alpha && !d->d_alpha) return fail(LDPC_ERR_INVALID, "decoder was created without alpha");
    DeviceGuard guard(d->g->device);
    if (!guard.ok) return fail(LDPC_ERR_CUDA, "cannot select CUDA device %d", d->g->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const ldpc_graph* g = d->g;
    const int64_t B = tc.B, Bp = tc.Bp;
    const size_t E = (size_t)std::max<int64_t>(g->E, 1);
    const size_t slice = E * (size_t)Bp;
    LAUNCH(K_OTHER, launch_pack(LDPC_F32, grad_posterior, tc.g_post, B, Bp, g->n, nullptr, nullptr, nullptr, d->T, stream));
    // weight gradients: tables of a few columns (degree-shared weights) are accumulated in `kParts` spread copies and
    // folded at the end -- every warp adding to the same few addresses serialises in L2 (TrainBwd in ldpc_internal.h)
    constexpr int kParts = 64;
    auto parts_for = [&](int cols) { return (cols > 0 && cols <= 4096) ? kParts : 1; };
    const int beta_parts = grad_beta ? parts_for(d->n_beta) : 1, alpha_parts = grad_alpha ? parts_for(d->n_alpha) : 1;
    const size_t beta_count = (size_t)d->T * d->n_beta, alpha_count = (size_t)d->T * d->n_alpha;
    const size_t part_floats = (beta_parts > 1 ? beta_parts * beta_count : 0) + (alpha_parts > 1 ? alpha_parts * alpha_count : 0);
    if (part_floats > tc.part_cap) {
        cudaFree(tc.g_part);
        tc.g_part = nullptr;
        tc.part_cap = 0;
        CU(cudaMalloc((void**)&tc.g_part, part_floats * sizeof(float)));
        tc.part_cap = part_floats;
    }
    float* const beta_acc = beta_parts > 1 ? tc.g_part : grad_beta;
    float* const alpha_acc = alpha_parts > 1 ? tc.g_part + (beta_parts > 1 ? beta_parts * beta_count : 0) : grad_alpha;
    if (grad_beta) CU(cudaMemsetAsync(beta_acc, 0, (size_t)beta_parts * beta_count * sizeof(float), stream));
    if (grad_alpha) CU(cudaMemsetAsync(alpha_acc, 0, (size_t)alpha_parts * alpha_count * sizeof(float), stream));
    TrainBwd p{};
    p.B = B;
    p.Bp = Bp;
    p.iters = tc.ws.iters;
    p.cn_items = g->cn[0].d;
    p.n_cn_items = (int)g->cn[0].items.size();
    p.vn_items = g->vn[0].d;
    p.n_vn_items = (int)g->vn[0].items.size();
    p.slot_var = g->d_slot_var;
    p.vslots = g->d_vslots;
    p.vpos_var = g->d_vpos_var;
    p.bidx = d->d_bidx;
    p.beta_per_edge = d->beta_per_edge;
    p.beta = static_cast<const float*>(d->d_beta);
    p.n_beta = d->n_beta;
    p.beta_const = 1.f;
    p.aidx = d->d_aidx;
    p.alpha = static_cast<const float*>(d->d_alpha);
    p.n_alpha = d->n_alpha;
    p.llrT = static_cast<const float*>(tc.ws.llrT);
    p.g_post = tc.g_post;
    p.g_v2c = tc.g_v2c;
    p.g_c2v = tc.g_c2v;
    p.g_beta = grad_beta ? beta_acc : nullptr;
    p.g_alpha = grad_alpha ? alpha_acc : nullptr;
    p.beta_parts = beta_parts;
    p.alpha_parts = alpha_parts;
    p.T = d->T;
    p.force_general = d->train_general;
    for (int t = d->T - 1; t >= 0; --t) {
        p.v2c_t = tc.v2c_hist + (size_t)t * slice;
        p.c2v_t = tc.c2v_hist + (size_t)t * slice;
        LAUNCH(K_OTHER, launch_train_bwd_vn(p, t, stream));
        LAUNCH(K_OTHER, launch_train_bwd_cn(p, t, stream));
    }
    if (grad_beta && beta_parts > 1) LAUNCH(K_OTHER, launch_train_fold(beta_acc, grad_beta, beta_parts, (int64_t)beta_count, stream));
    if (grad_alpha && alpha_parts > 1) LAUNCH(K_OTHER, launch_train_fold(alpha_acc, grad_alpha, alpha_parts, (int64_t)alpha_count, stream));
    return LDPC_OK;
}

// =================================================================================================
// Monte-Carlo leg
// =================================================================================================
extern "C" int ldpc_awgn_llr(int device, int32_t n, int64_t B, uint64_t frame0, uint64_t seed, float snr_db,
                             int32_t llr_sign, const uint8_t* codeword, float* llr_out, void* stream) {
    if (n < 1 || B < 1 || !llr_out) return fail(LDPC_ERR_INVALID, "bad arguments");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(LDPC_ERR_CUDA, "cannot select CUDA device %d", device);
    cudaError_t e = launch_awgn(LDPC_F32, 1, llr_out, n, B, B, frame0, seed, snr_db, llr_sign, codeword, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(LDPC_ERR_CUDA, "awgn launch: %s", cudaGetErrorString(e));
    return LDPC_OK;
}

extern "C" int ldpc_mc_round(ldpc_decoder* d, float snr_db, int32_t llr_sign, uint64_t seed, uint64_t frame0, int64_t B,
                             const uint8_t* codeword, int64_t* counters, int32_t* frame_bit_errors,
                             int32_t* frame_iterations, void* stream_) {
    if (!d || !counters) return fail(LDPC_ERR_INVALID, "NULL argument");
    if (B < 1) return fail(LDPC_ERR_INVALID, "B must be >= 1");
    DeviceGuard guard(d->g->device);
    if (!guard.ok) return fail(LDPC_ERR_CUDA, "cannot select CUDA device %d", d->g->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const int64_t Bp = pad_frames(B);
    int rc = ctx_enter(d->cx[0], stream);
    if (!rc) rc = ws_ensure(d, d->cx[0].root, Bp);
    if (rc) return rc;
    d->prof.frames_padded = Bp;
    Workspace& ws = d->cx[0].root;
    LAUNCH(K_OTHER, launch_awgn(d->dtype, 0, ws.llrT, d->g->n, B, Bp, frame0, seed, snr_db, llr_sign, codeword, stream));
    OutSpec o;
    o.count = true;
    o.codeword = codeword;
    o.counters = counters;
    o.frame_bit_errors = frame_bit_errors;
    o.frame_iters = frame_iterations;
    rc = decode_resident(d, d->cx[0], B, Bp, o, stream);
    if (!rc) rc = ctx_leave(d->cx[0], stream);
    return rc;
}

extern "C" int ldpc_count_errors(int device, int32_t n, int64_t B, const uint8_t* bits, const uint8_t* codeword,
                                 const int32_t* iterations, int64_t* counters, int32_t* frame_bit_errors, void* stream) {
    if (n < 1 || B < 1 || !bits || !counters) return fail(LDPC_ERR_INVALID, "bad arguments");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(LDPC_ERR_CUDA, "cannot select CUDA device %d", device);
    cudaError_t e = launch_count_bits(bits, n, B, codeword, iterations, counters, frame_bit_errors, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(LDPC_ERR_CUDA, "count launch: %s", cudaGetErrorString(e));
    return LDPC_OK;
}

// =================================================================================================
// Instrumentation
// =================================================================================================
extern "C" int ldpc_decoder_profile_mode(ldpc_decoder* d, int32_t mode) {
    if (!d || mode < 0 || mode > 1) return fail(LDPC_ERR_INVALID, "bad arguments");
    d->prof_mode = mode;
    return LDPC_OK;
}

extern "C" int ldpc_decoder_profile_read(ldpc_decoder* d, ldpc_profile* out, int32_t reset) {
    if (!d || !out) return fail(LDPC_ERR_INVALID, "NULL argument");
    DeviceGuard guard(d->g->device);
    if (!d->pending.empty()) {
        CU(cudaDeviceSynchronize());
        for (auto& p : d->pending) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
                if (p.kind == K_CN) d->prof.cn_ms += ms;
                else if (p.kind == K_VN) d->prof.vn_ms += ms;
                else d->prof.other_ms += ms;
            }
        }
        d->pending.clear();
        d->ev_next = 0;
    }
    d->prof.compactions = d->stat_compactions;
    d->prof.early_exits = d->stat_early_exits;
    d->prof.graph_replays = d->stat_graph_replays;
    d->prof.small_decodes = d->stat_small;
    d->prof.resident_decodes = d->stat_resident;
    *out = d->prof;
    if (reset) {
        d->stat_compactions = d->stat_early_exits = d->stat_graph_replays = d->stat_small = d->stat_resident = 0;
        int64_t fp = d->prof.frames_padded;
        d->prof = ldpc_profile{};
        d->prof.frames_padded = fp;
    }
    return LDPC_OK;
}
