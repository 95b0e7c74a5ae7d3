// Python bindings (pybind11, no torch C++ dependency: tensors cross the boundary as raw device
// pointers + CUDA stream handles) and the symmetric peer-memory allocator.
//
// The allocator is the native replacement of the reference's mpi4py communicator bootstrap
// (VGG/allreducer.py:219-220): every rank cudaMalloc's a block, exports a CUDA IPC handle, and
// maps every peer's block into its own address space, after which kernels address peer memory
// with plain ld/st/red/cp.async.bulk over NVLink 5 / NVSwitch.
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>
#include <cuda_runtime.h>

#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "oktopk.cuh"

namespace py = pybind11;
using namespace okt;

static void ck(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}
template <class T> static T* P_(uint64_t p) { return reinterpret_cast<T*>(p); }
static cudaStream_t S_(uint64_t s) { return reinterpret_cast<cudaStream_t>(s); }

// ------------------------------------------------------------------------------------------- symmetric memory
static py::tuple symm_alloc(size_t nbytes) {
    void* p = nullptr;
    ck(cudaMalloc(&p, nbytes), "cudaMalloc(symm)");
    ck(cudaMemset(p, 0, nbytes), "cudaMemset(symm)");
    cudaIpcMemHandle_t h;
    ck(cudaIpcGetMemHandle(&h, p), "cudaIpcGetMemHandle");
    return py::make_tuple((uint64_t)p, py::bytes(reinterpret_cast<const char*>(&h), sizeof(h)));
}
static uint64_t symm_open(const std::string& handle) {
    if (handle.size() != sizeof(cudaIpcMemHandle_t)) throw std::runtime_error("bad IPC handle size");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle.data(), sizeof(h));
    void* p = nullptr;
    ck(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
    return (uint64_t)p;
}
static void symm_close(uint64_t p) { ck(cudaIpcCloseMemHandle(P_<void>(p)), "cudaIpcCloseMemHandle"); }
static void symm_free(uint64_t p) { ck(cudaFree(P_<void>(p)), "cudaFree(symm)"); }
static uint64_t dev_alloc_zero(size_t nbytes) {
    void* p = nullptr;
    ck(cudaMalloc(&p, nbytes), "cudaMalloc");
    ck(cudaMemset(p, 0, nbytes), "cudaMemset");
    return (uint64_t)p;
}
static void dev_free(uint64_t p) {
    if (p) ck(cudaFree(P_<void>(p)), "cudaFree");
}
static void memset_async(uint64_t p, int value, size_t nbytes, uint64_t stream) {
    ck(cudaMemsetAsync(P_<void>(p), value, nbytes, S_(stream)), "cudaMemsetAsync");
}
static int can_access_peer(int dev, int peer) {
    int ok = 0;
    ck(cudaDeviceCanAccessPeer(&ok, dev, peer), "cudaDeviceCanAccessPeer");
    return ok;
}

// ------------------------------------------------------------------------------------------- state access
static size_t state_bytes() { return sizeof(OktState); }

static py::dict layout_info(int P, int cap, int gcap) {
    SymmLayout L = make_layout(P, cap, gcap);
    py::dict d;
    d["total"] = L.total; d["rs_mbox"] = L.rs_mbox; d["rs_thr"] = L.rs_thr; d["ag_mbox"] = L.ag_mbox;
    d["cut_mbox"] = L.cut_mbox; d["cut_data"] = L.cut_data; d["send_idx"] = L.send_idx; d["send_val"] = L.send_val;
    d["gat_idx"] = L.gat_idx; d["gat_val"] = L.gat_val; d["cap"] = L.cap; d["gcap"] = L.gcap;
    d["chunk"] = kChunk; d["maxp"] = OKT_MAXP; d["threads"] = kThreads;
    return d;
}

// Synchronous (stream-ordered copy + sync): observability only, never on the hot path.
static py::dict read_state(uint64_t st, int P, uint64_t stream) {
    static thread_local std::vector<char> host(offsetof(OktState, hist));
    ck(cudaMemcpyAsync(host.data(), P_<void>(st), host.size(), cudaMemcpyDeviceToHost, S_(stream)), "read_state");
    ck(cudaStreamSynchronize(S_(stream)), "read_state sync");
    const OktState* s = reinterpret_cast<const OktState*>(host.data());
    py::dict d;
    d["local_thr"] = s->local_thr; d["local_thr_used"] = s->local_thr_used; d["global_thr"] = s->global_thr;
    d["epoch"] = s->epoch;
    std::vector<int> e(s->edges, s->edges + P + 1);
    d["edges"] = e;
    d["local_count"] = s->stat_local_count; d["global_count"] = s->stat_global_count;
    d["recv_total"] = s->stat_recv_total; d["gather_total"] = s->stat_gather_total;
    d["overflow_send"] = s->stat_overflow_send; d["overflow_gather"] = s->stat_overflow_gather;
    d["fault"] = s->fault;
    // phase durations of the last fused call in microseconds: pack, reduce-scatter, global select, allgather+finalise
    auto us = [&](int a, int b) { return s->t_phase[b] >= s->t_phase[a] ? (double)(s->t_phase[b] - s->t_phase[a]) * 1e-3 : 0.0; };
    py::dict ph;
    ph["pack"] = us(0, 1); ph["reduce"] = us(1, 2); ph["gselect"] = us(2, 3); ph["final"] = us(3, 4); ph["total"] = us(0, 4);
    d["phase_us"] = ph;
    return d;
}

static void write_state(uint64_t st, float local_thr, float global_thr, const std::vector<int>& edges, uint64_t stream) {
    std::vector<char> host(offsetof(OktState, hist));
    ck(cudaMemcpyAsync(host.data(), P_<void>(st), host.size(), cudaMemcpyDeviceToHost, S_(stream)), "write_state rd");
    ck(cudaStreamSynchronize(S_(stream)), "write_state sync");
    OktState* s = reinterpret_cast<OktState*>(host.data());
    s->local_thr = local_thr;
    s->global_thr = global_thr;
    if (!edges.empty()) {
        if (edges.size() > OKT_MAXP + 1) throw std::runtime_error("too many region edges");
        for (size_t i = 0; i < edges.size(); ++i) s->edges[i] = edges[i];
    }
    // only the plain-data head is written back (bar / epoch / cursors are kernel-owned and unchanged here)
    ck(cudaMemcpyAsync(P_<void>(st), host.data(), host.size(), cudaMemcpyHostToDevice, S_(stream)), "write_state wr");
    ck(cudaStreamSynchronize(S_(stream)), "write_state sync2");
}

// ------------------------------------------------------------------------------------------- launchers
static void fill_peers(char** dst, const std::vector<uint64_t>& peers) {
    if (peers.size() > OKT_MAXP) throw std::runtime_error("world larger than OKT_MAXP");
    for (int i = 0; i < OKT_MAXP; ++i) dst[i] = nullptr;
    for (size_t i = 0; i < peers.size(); ++i) dst[i] = P_<char>(peers[i]);
}

static void oktopk_run(uint64_t g, uint64_t res, uint64_t st, const std::vector<uint64_t>& peers, int n, int rank,
                       int k, int cap, int gcap, py::dict o, int grid, uint64_t stream) {
    OktParams p;
    std::memset(&p, 0, sizeof(p));
    p.g = P_<float>(g); p.res = P_<float>(res); p.st = P_<OktState>(st);
    if (!o.contains("cand") || !o.contains("ccap")) throw std::runtime_error("oktopk_run: candidate scratch missing");
    p.cand = P_<int>(o["cand"].cast<uint64_t>()); p.ccap = o["ccap"].cast<int>();
    p.cand_mode = o.contains("cand_mode") ? o["cand_mode"].cast<int>() : 1;
    fill_peers(p.peers, peers);
    p.P = (int)peers.size(); p.rank = rank; p.n = n; p.k = k;
    p.L = make_layout(p.P, cap, gcap);
    auto geti = [&](const char* key, int dflt) { return o.contains(key) ? o[key].cast<int>() : dflt; };
    auto getf = [&](const char* key, double dflt) { return o.contains(key) ? o[key].cast<double>() : dflt; };
    p.exact_local = geti("exact_local", 0);
    p.repartition = geti("repartition", 0);
    p.uniform_regions = geti("uniform_regions", 0);
    p.residual_mode = geti("residual_mode", RES_OKTOPK);
    p.global_mode = geti("global_mode", GLB_THRESHOLD);
    p.deterministic = geti("deterministic", 0);
    p.pull_tma = geti("pull_tma", 1);
    p.phase_begin = geti("phase_begin", 0);
    p.phase_end = geti("phase_end", PH_END);
    p.guard_loops = geti("guard_loops", 0);
    if (p.guard_loops > kGuardMax - 1) p.guard_loops = kGuardMax - 1;
    p.guard_limit = geti("guard_limit", 0);
    p.guard_factor = (float)getf("guard_factor", 1.03);
    p.l_low_cnt = getf("l_low_cnt", 0.0); p.l_high_cnt = getf("l_high_cnt", 1e30);
    p.l_factor = (float)getf("l_factor", 1.012);
    p.g_low_cnt = getf("g_low_cnt", 0.0); p.g_high_cnt = getf("g_high_cnt", 1e30);
    p.g_inc = (float)getf("g_inc", 1.008); p.g_dec = (float)getf("g_dec", 1.008);
    p.prefilter = (float)getf("prefilter", 0.8);
    p.timeout_ns = (unsigned long long)(getf("timeout_s", 0.0) * 1e9);
    if (geti("split_phases", 0)) {
        // ablation / debugging: one launch per phase instead of the single persistent kernel
        for (int ph = p.phase_begin; ph < p.phase_end; ++ph) {
            OktParams q = p;
            q.phase_begin = ph; q.phase_end = ph + 1;
            ck(launch_oktopk(q, grid, S_(stream)), "oktopk phase launch");
        }
    } else {
        ck(launch_oktopk(p, grid, S_(stream)), "oktopk fused launch");
    }
}

static void gather_run(uint64_t g, uint64_t res, uint64_t st, const std::vector<uint64_t>& peers, int n, int rank,
                       int k, int cap, int gcap, py::dict o, int grid, uint64_t stream) {
    GatherParams p;
    std::memset(&p, 0, sizeof(p));
    p.g = P_<float>(g); p.res = P_<float>(res); p.st = P_<OktState>(st);
    fill_peers(p.peers, peers);
    p.P = (int)peers.size(); p.rank = rank; p.n = n; p.k = k;
    p.L = make_layout(p.P, cap, gcap);
    p.select_mode = o["select_mode"].cast<int>();
    p.exact_now = o.contains("exact_now") ? o["exact_now"].cast<int>() : 0;
    p.gauss_mode = o.contains("gauss_mode") ? o["gauss_mode"].cast<int>() : 0;
    p.gauss_loops = o.contains("gauss_loops") ? o["gauss_loops"].cast<int>() : 20;
    p.gauss_factor = o.contains("gauss_factor") ? (float)o["gauss_factor"].cast<double>() : 1.02f;
    p.density = (float)o["density"].cast<double>();
    p.pull_tma = o.contains("pull_tma") ? o["pull_tma"].cast<int>() : 1;
    p.timeout_ns = o.contains("timeout_s") ? (unsigned long long)(o["timeout_s"].cast<double>() * 1e9) : 0ULL;
    ck(launch_gather_scheme(p, grid, S_(stream)), "gather scheme launch");
}

static void dense_run(const std::vector<uint64_t>& bufs, const std::vector<uint64_t>& flags, uint64_t epoch, int n,
                      int rank, int grid, uint64_t stream, uint64_t st, double timeout_s) {
    DenseParams p;
    std::memset(&p, 0, sizeof(p));
    if (bufs.size() > OKT_MAXP || bufs.size() != flags.size()) throw std::runtime_error("bad peer tables");
    for (size_t i = 0; i < bufs.size(); ++i) { p.bufs[i] = P_<float>(bufs[i]); p.flags[i] = P_<uint64_t>(flags[i]); }
    p.epoch = P_<unsigned long long>(epoch);
    p.n = n; p.P = (int)bufs.size(); p.rank = rank; p.scale = 1.0f / (float)p.P;
    p.fault = st ? &P_<OktState>(st)->fault : nullptr;
    p.timeout_ns = (unsigned long long)(timeout_s * 1e9);
    ck(launch_dense_allreduce(p, grid, S_(stream)), "dense allreduce launch");
}

static void kth_abs(uint64_t x, int n, int k, uint64_t st, uint64_t out, int grid, uint64_t stream) {
    ck(launch_kth_abs(P_<float>(x), n, k, P_<OktState>(st), P_<float>(out), grid, S_(stream)), "kth_abs launch");
}

static void fused_sgd(uint64_t p, uint64_t g, uint64_t mom, int n, double lr, double momentum, double dampening,
                      double wd, int nesterov, int first, int zero_grad, double grad_scale, uint64_t stream,
                      uint64_t lr_ptr) {
    ck(launch_fused_sgd(P_<float>(p), P_<float>(g), P_<float>(mom), n, (float)lr, (float)momentum, (float)dampening,
                        (float)wd, nesterov, first, zero_grad, (float)grad_scale, P_<float>(lr_ptr), S_(stream)),
       "fused_sgd");
}
static void fused_bert_adam(uint64_t p, uint64_t g, uint64_t m, uint64_t v, int n, double lr, double b1, double b2,
                            double eps, double wd, int zero_grad, uint64_t stream, uint64_t lr_ptr) {
    ck(launch_fused_bert_adam(P_<float>(p), P_<float>(g), P_<float>(m), P_<float>(v), n, (float)lr, (float)b1,
                              (float)b2, (float)eps, (float)wd, zero_grad, P_<float>(lr_ptr), S_(stream)),
       "fused_bert_adam");
}
static void momentum_correct(uint64_t g, uint64_t buf, int n, double momentum, uint64_t stream) {
    ck(launch_momentum_correct(P_<float>(g), P_<float>(buf), n, (float)momentum, S_(stream)), "momentum_correct");
}
static void clip_by_norm(uint64_t x, int n, uint64_t scratch, double max_norm, uint64_t stream) {
    ck(launch_l2norm_sq(P_<float>(x), n, P_<float>(scratch), S_(stream)), "l2norm");
    ck(launch_scale(P_<float>(x), n, P_<float>(scratch), (float)max_norm, S_(stream)), "clip scale");
}

PYBIND11_MODULE(_C, m) {
    m.doc() = "oktopk_b200 native extension (sm_100a kernels + symmetric peer memory)";
    m.def("symm_alloc", &symm_alloc);
    m.def("symm_open", &symm_open);
    m.def("symm_close", &symm_close);
    m.def("symm_free", &symm_free);
    m.def("dev_alloc_zero", &dev_alloc_zero);
    m.def("dev_free", &dev_free);
    m.def("memset_async", &memset_async);
    m.def("can_access_peer", &can_access_peer);
    m.def("state_bytes", &state_bytes);
    m.def("layout_info", &layout_info);
    m.def("read_state", &read_state);
    m.def("write_state", &write_state);
    m.def("max_coop_grid", &okt_max_coop_grid);
    m.def("oktopk_run", &oktopk_run);
    m.def("gather_run", &gather_run);
    m.def("dense_run", &dense_run, py::arg("bufs"), py::arg("flags"), py::arg("epoch"), py::arg("n"), py::arg("rank"),
          py::arg("grid"), py::arg("stream"), py::arg("st") = 0, py::arg("timeout_s") = 0.0);
    m.def("clear_fault", [](uint64_t st, uint64_t stream) {
        ck(cudaMemsetAsync(&P_<OktState>(st)->fault, 0, sizeof(int), S_(stream)), "clear_fault");
    });
    m.def("kth_abs", &kth_abs);
    m.def("fused_sgd", &fused_sgd, py::arg("p"), py::arg("g"), py::arg("mom"), py::arg("n"), py::arg("lr"),
          py::arg("momentum"), py::arg("dampening"), py::arg("wd"), py::arg("nesterov"), py::arg("first"),
          py::arg("zero_grad"), py::arg("grad_scale"), py::arg("stream"), py::arg("lr_ptr") = 0);
    m.def("fused_bert_adam", &fused_bert_adam, py::arg("p"), py::arg("g"), py::arg("m"), py::arg("v"), py::arg("n"),
          py::arg("lr"), py::arg("b1"), py::arg("b2"), py::arg("eps"), py::arg("wd"), py::arg("zero_grad"),
          py::arg("stream"), py::arg("lr_ptr") = 0);
    m.def("momentum_correct", &momentum_correct);
    m.def("clip_by_norm", &clip_by_norm);
    m.attr("MAXP") = OKT_MAXP;
    m.attr("CHUNK") = kChunk;
}
