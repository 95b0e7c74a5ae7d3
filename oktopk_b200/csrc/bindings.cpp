// Python bindings (pybind11, no torch C++ dependency: tensors cross the boundary as raw device
// pointers + CUDA stream handles) and the symmetric peer-memory allocator.
//
// The allocator is the native replacement of the reference's mpi4py communicator bootstrap
// (VGG/allreducer.py:219-220): every rank cudaMalloc's a block, exports a CUDA IPC handle, and
// maps every peer's block into its own address space, after which kernels address peer memory
// with plain ld/st/red/cp.async.bulk over NVLink 5 / NVSwitch.
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>
#include <cuda.h>            // driver-API TYPES only: the entry points are resolved at run time (cudaGetDriverEntryPoint),
#include <cuda_runtime.h>    // so the module imports on a box without libcuda (the CPU build check)

#include <unistd.h>

#include <algorithm>
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


// ------------------------------------------------------------------------------------------- VMM + NVSwitch multicast
// The NVLS dense path needs the bucket mapped through a MULTICAST object (multimem.ld_reduce / multimem.st address
// all P copies at once), which CUDA IPC handles cannot provide: the block is then allocated with the virtual-memory
// API (cuMemCreate, exported as a POSIX file descriptor that Python passes to the peers over a unix socket),
// mapped by every peer (unicast, as before) and bound to a multicast object created by rank 0.
// Driver entry points are looked up at run time so that this module has no link-time dependency on libcuda.
namespace drv {
template <class F> static F sym(const char* name) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || fn == nullptr || q != cudaDriverEntryPointSuccess)
        throw std::runtime_error(std::string("driver entry point not available: ") + name);
    return reinterpret_cast<F>(fn);
}
static void ckd(CUresult r, const char* what) {
    if (r != CUDA_SUCCESS) throw std::runtime_error(std::string(what) + ": CUresult " + std::to_string((int)r));
}
}  // namespace drv

static CUmemAllocationProp vmm_prop(int dev) {
    CUmemAllocationProp prop;
    std::memset(&prop, 0, sizeof(prop));
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = dev;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    return prop;
}

// {vmm: 0/1, posix_fd: 0/1, multicast: 0/1, granularity, mc_granularity}
static py::dict vmm_probe(int dev, int ndev) {
    py::dict d;
    d["vmm"] = 0; d["posix_fd"] = 0; d["multicast"] = 0; d["granularity"] = 0; d["mc_granularity"] = 0;
    try {
        auto getattr_ = drv::sym<CUresult (*)(int*, CUdevice_attribute, CUdevice)>("cuDeviceGetAttribute");
        int v = 0;
        if (getattr_(&v, CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED, dev) == CUDA_SUCCESS) d["vmm"] = v;
        if (getattr_(&v, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, dev) == CUDA_SUCCESS) d["posix_fd"] = v;
        if (getattr_(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) == CUDA_SUCCESS) d["multicast"] = v;
        auto gran = drv::sym<CUresult (*)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags)>("cuMemGetAllocationGranularity");
        CUmemAllocationProp prop = vmm_prop(dev);
        size_t g = 0;
        if (gran(&g, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS) d["granularity"] = g;
        if (d["multicast"].cast<int>() && ndev > 1) {
            auto mgran = drv::sym<CUresult (*)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags)>("cuMulticastGetGranularity");
            CUmulticastObjectProp mp;
            std::memset(&mp, 0, sizeof(mp));
            mp.numDevices = ndev; mp.size = 1 << 21; mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
            size_t mg = 0;
            if (mgran(&mg, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS) d["mc_granularity"] = mg;
        }
    } catch (const std::exception& e) {
        d["error"] = std::string(e.what());
    }
    return d;
}

// physical allocation on `dev`, exported as a POSIX fd: (handle, fd)
static py::tuple vmm_create(size_t nbytes, int dev) {
    auto create = drv::sym<CUresult (*)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long)>("cuMemCreate");
    auto exp = drv::sym<CUresult (*)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long)>("cuMemExportToShareableHandle");
    CUmemAllocationProp prop = vmm_prop(dev);
    CUmemGenericAllocationHandle h = 0;
    drv::ckd(create(&h, nbytes, &prop, 0), "cuMemCreate");
    int fd = -1;
    drv::ckd(exp(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0), "cuMemExportToShareableHandle");
    return py::make_tuple((uint64_t)h, fd);
}
static uint64_t vmm_import(int fd) {
    auto imp = drv::sym<CUresult (*)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType)>("cuMemImportFromShareableHandle");
    CUmemGenericAllocationHandle h = 0;
    drv::ckd(imp(&h, reinterpret_cast<void*>(static_cast<uintptr_t>(fd)), CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR), "cuMemImportFromShareableHandle");
    return (uint64_t)h;
}
// map a (memory or multicast) handle into this process with read/write access for `dev`
static uint64_t vmm_map(uint64_t handle, size_t nbytes, int dev, size_t align) {
    auto reserve = drv::sym<CUresult (*)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long)>("cuMemAddressReserve");
    auto map = drv::sym<CUresult (*)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long)>("cuMemMap");
    auto access = drv::sym<CUresult (*)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t)>("cuMemSetAccess");
    CUdeviceptr va = 0;
    drv::ckd(reserve(&va, nbytes, align, 0, 0), "cuMemAddressReserve");
    drv::ckd(map(va, nbytes, 0, (CUmemGenericAllocationHandle)handle, 0), "cuMemMap");
    CUmemAccessDesc ad;
    std::memset(&ad, 0, sizeof(ad));
    ad.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    ad.location.id = dev;
    ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    drv::ckd(access(va, nbytes, &ad, 1), "cuMemSetAccess");
    return (uint64_t)va;
}
static void vmm_unmap(uint64_t va, size_t nbytes) {
    auto unmap = drv::sym<CUresult (*)(CUdeviceptr, size_t)>("cuMemUnmap");
    auto afree = drv::sym<CUresult (*)(CUdeviceptr, size_t)>("cuMemAddressFree");
    unmap((CUdeviceptr)va, nbytes);
    afree((CUdeviceptr)va, nbytes);
}
static void vmm_release(uint64_t handle) {
    auto rel = drv::sym<CUresult (*)(CUmemGenericAllocationHandle)>("cuMemRelease");
    rel((CUmemGenericAllocationHandle)handle);
}
// multicast object over `ndev` devices: (handle, fd).  Created by one rank, imported (vmm_import) by the others.
static py::tuple mc_create(size_t nbytes, int ndev) {
    auto create = drv::sym<CUresult (*)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*)>("cuMulticastCreate");
    auto exp = drv::sym<CUresult (*)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long)>("cuMemExportToShareableHandle");
    CUmulticastObjectProp mp;
    std::memset(&mp, 0, sizeof(mp));
    mp.numDevices = ndev; mp.size = nbytes; mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    CUmemGenericAllocationHandle h = 0;
    drv::ckd(create(&h, &mp), "cuMulticastCreate");
    int fd = -1;
    drv::ckd(exp(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0), "cuMemExportToShareableHandle(mc)");
    return py::make_tuple((uint64_t)h, fd);
}
static void mc_add_device(uint64_t mc, int dev) {
    auto add = drv::sym<CUresult (*)(CUmemGenericAllocationHandle, CUdevice)>("cuMulticastAddDevice");
    drv::ckd(add((CUmemGenericAllocationHandle)mc, dev), "cuMulticastAddDevice");
}
static void mc_bind(uint64_t mc, uint64_t mem, size_t nbytes) {
    auto bind = drv::sym<CUresult (*)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long)>("cuMulticastBindMem");
    drv::ckd(bind((CUmemGenericAllocationHandle)mc, 0, (CUmemGenericAllocationHandle)mem, 0, nbytes, 0), "cuMulticastBindMem");
}

// ------------------------------------------------------------------------------------------- state access
static size_t state_bytes() { return sizeof(OktState); }

static py::dict layout_info(int P, int n, int cap, int gcap) {
    SymmLayout L = make_layout(P, n, cap, gcap);
    py::dict d;
    d["total"] = L.total; d["rs_mbox"] = L.rs_mbox; d["rs_thr"] = L.rs_thr; d["ag_mbox"] = L.ag_mbox;
    d["cut_mbox"] = L.cut_mbox; d["cut_data"] = L.cut_data; d["send_idx"] = L.send_idx; d["send_val"] = L.send_val;
    d["gat_idx"] = L.gat_idx; d["gat_val"] = L.gat_val; d["cap"] = L.cap; d["gcap"] = L.gcap; d["scap"] = L.scap;
    d["done_mbox"] = L.done_mbox; d["tree_mbox"] = L.tree_mbox;
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
    d["redo"] = s->stat_redo; d["dense_fallback"] = s->stat_dense_fallback; d["pack_thr"] = s->pack_thr;
    d["cum_overflow_send"] = s->cum_overflow_send; d["cum_overflow_gather"] = s->cum_overflow_gather;
    d["cum_redo"] = s->cum_redo;
    d["fault"] = s->fault;
    // phase durations of the last fused call in microseconds: pack, reduce-scatter, global select, allgather+finalise
    auto us = [&](int a, int b) { return s->t_phase[b] >= s->t_phase[a] ? (double)(s->t_phase[b] - s->t_phase[a]) * 1e-3 : 0.0; };
    py::dict ph;
    ph["local"] = us(5, 0); ph["pack"] = us(0, 1); ph["wait_rs"] = us(1, 6); ph["reduce"] = us(6, 2); ph["gselect"] = us(2, 3);
    ph["wait_ag"] = us(3, 7); ph["final"] = us(7, 4); ph["total"] = us(5, 4);
    d["phase_us"] = ph;
    return d;
}

// The per-call history ring (one TraceRec per fused call, newest kTraceLen calls): --trace / settings.PROFILING.
static py::list read_trace(uint64_t st, uint64_t stream) {
    std::vector<TraceRec> host(kTraceLen);
    ck(cudaMemcpyAsync(host.data(), reinterpret_cast<char*>(P_<OktState>(st)) + offsetof(OktState, trace),
                       sizeof(TraceRec) * kTraceLen, cudaMemcpyDeviceToHost, S_(stream)), "read_trace");
    ck(cudaStreamSynchronize(S_(stream)), "read_trace sync");
    py::list out;
    for (const TraceRec& t : host) {
        if (t.epoch == 0) continue;
        py::dict d;
        d["epoch"] = t.epoch; d["local_count"] = t.local_count; d["global_count"] = t.global_count;
        d["recv_total"] = t.recv_total; d["gather_total"] = t.gather_total; d["overflow_send"] = t.overflow_send;
        d["overflow_gather"] = t.overflow_gather; d["redo"] = t.redo; d["local_thr"] = t.local_thr;
        d["global_thr"] = t.global_thr; d["us_local"] = t.us_local; d["us_pack"] = t.us_pack;
        d["us_wait_rs"] = t.us_wait_rs; d["us_reduce"] = t.us_reduce; d["us_gselect"] = t.us_gselect;
        d["us_wait_ag"] = t.us_wait_ag; d["us_final"] = t.us_final; d["t_begin"] = t.t_begin;
        out.append(d);
    }
    return out;
}

// A host-mapped pinned int: kernels mirror their fault code into it, the host polls it at every step for free.
static py::tuple host_flag_alloc() {
    int* h = nullptr;
    ck(cudaHostAlloc(reinterpret_cast<void**>(&h), sizeof(int) * 16, cudaHostAllocMapped), "cudaHostAlloc(flag)");
    for (int i = 0; i < 16; ++i) h[i] = 0;
    void* d = nullptr;
    ck(cudaHostGetDevicePointer(&d, h, 0), "cudaHostGetDevicePointer");
    return py::make_tuple((uint64_t)h, (uint64_t)d);
}
static void host_flag_free(uint64_t h) { if (h) cudaFreeHost(P_<void>(h)); }
static int host_flag_read(uint64_t h) { return h ? *reinterpret_cast<volatile int*>(h) : 0; }
static void host_flag_clear(uint64_t h) { if (h) *reinterpret_cast<volatile int*>(h) = 0; }

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
    p.L = make_layout(p.P, n, cap, gcap);
    auto geti = [&](const char* key, int dflt) { return o.contains(key) ? o[key].cast<int>() : dflt; };
    auto getf = [&](const char* key, double dflt) { return o.contains(key) ? o[key].cast<double>() : dflt; };
    for (int i = 0; i < OKT_MAXP; ++i) p.peer_g[i] = nullptr;
    if (o.contains("peer_g")) {
        auto pg = o["peer_g"].cast<std::vector<uint64_t>>();
        for (size_t i = 0; i < pg.size() && i < OKT_MAXP; ++i) p.peer_g[i] = P_<float>(pg[i]);
    }
    p.max_redo = geti("max_redo", 12);
    p.redo_factor = (float)getf("redo_factor", 1.5);
    p.dense_nnz_limit = geti("dense_nnz_limit", 0);
    p.host_fault = o.contains("host_fault") ? P_<int>(o["host_fault"].cast<uint64_t>()) : nullptr;
    p.trace = geti("trace", 0);
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
    if (p.guard_loops > kGuardFineMax) p.guard_loops = kGuardFineMax;
    p.cap_limit = geti("cap_limit", 0);
    p.cap_rungs = geti("cap_rungs", 40);
    p.cap_factor = (float)getf("cap_factor", 1.19);
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
    p.L = make_layout(p.P, n, cap, gcap);
    p.reselect = o.contains("reselect") ? o["reselect"].cast<int>() : 0;
    p.clip_max_norm = o.contains("clip_max_norm") ? (float)o["clip_max_norm"].cast<double>() : 0.f;
    p.bitmap = o.contains("bitmap") ? P_<unsigned>(o["bitmap"].cast<uint64_t>()) : nullptr;
    p.cand = o.contains("cand") ? P_<int>(o["cand"].cast<uint64_t>()) : nullptr;
    p.ccap = o.contains("ccap") ? o["ccap"].cast<int>() : 0;
    p.host_fault = o.contains("host_fault") ? P_<int>(o["host_fault"].cast<uint64_t>()) : nullptr;
    if (p.reselect && (p.bitmap == nullptr || p.cand == nullptr || p.ccap <= 0))
        throw std::runtime_error("gather_run: TopkA2 needs the bitmap and candidate scratch");
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

static void gtopk_run(uint64_t g, uint64_t res, uint64_t st, const std::vector<uint64_t>& peers, int n, int rank,
                      int k, int cap, int gcap, py::dict o, int grid, uint64_t stream) {
    TreeParams p;
    std::memset(&p, 0, sizeof(p));
    p.g = P_<float>(g); p.res = P_<float>(res); p.st = P_<OktState>(st);
    fill_peers(p.peers, peers);
    p.P = (int)peers.size(); p.rank = rank; p.n = n; p.k = k;
    if (p.P & (p.P - 1)) throw std::runtime_error("gTopk needs a power-of-two world size (VGG/allreducer.py:113)");
    p.L = make_layout(p.P, n, cap, gcap);
    p.pull_tma = o.contains("pull_tma") ? o["pull_tma"].cast<int>() : 1;
    p.timeout_ns = o.contains("timeout_s") ? (unsigned long long)(o["timeout_s"].cast<double>() * 1e9) : 0ULL;
    p.clip_max_norm = o.contains("clip_max_norm") ? (float)o["clip_max_norm"].cast<double>() : 0.f;
    p.bitmap = P_<unsigned>(o["bitmap"].cast<uint64_t>());
    p.cand = P_<int>(o["cand"].cast<uint64_t>());
    p.ccap = o["ccap"].cast<int>();
    p.sel_idx = P_<int>(o["sel_idx"].cast<uint64_t>());
    p.sel_val = P_<float>(o["sel_val"].cast<uint64_t>());
    p.selcap = o["selcap"].cast<int>();
    p.host_fault = o.contains("host_fault") ? P_<int>(o["host_fault"].cast<uint64_t>()) : nullptr;
    ck(launch_gtopk(p, grid, S_(stream)), "gtopk launch");
}

static void land_grads(const std::vector<uint64_t>& srcs, const std::vector<long long>& offs, const std::vector<int>& numels,
                       uint64_t bucket, uint64_t stream) {
    const size_t T = srcs.size();
    if (offs.size() != T || numels.size() != T) throw std::runtime_error("land_grads: table size mismatch");
    constexpr int per = 8192;                           // kLandPerCta
    for (size_t b = 0; b < T; b += kLandMax) {
        LandParams lp;
        std::memset(&lp, 0, sizeof(lp));
        const int cnt = (int)std::min<size_t>(kLandMax, T - b);
        int blk = 0;
        for (int i = 0; i < cnt; ++i) {
            lp.src[i] = P_<const float>(srcs[b + i]);
            lp.dst_off[i] = offs[b + i];
            lp.numel[i] = numels[b + i];
            lp.blk_begin[i] = blk;
            blk += std::max(1, (numels[b + i] + per - 1) / per);
        }
        lp.blk_begin[cnt] = blk;
        lp.count = cnt;
        ck(launch_land(lp, P_<float>(bucket), S_(stream)), "land_grads");
    }
}

static void dense_run(const std::vector<uint64_t>& bufs, const std::vector<uint64_t>& flags, uint64_t epoch, int n,
                      int rank, int grid, uint64_t stream, uint64_t st, double timeout_s, uint64_t mc,
                      uint64_t host_fault) {
    DenseParams p;
    std::memset(&p, 0, sizeof(p));
    if (bufs.size() > OKT_MAXP || bufs.size() != flags.size()) throw std::runtime_error("bad peer tables");
    for (size_t i = 0; i < bufs.size(); ++i) { p.bufs[i] = P_<float>(bufs[i]); p.flags[i] = P_<uint64_t>(flags[i]); }
    p.epoch = P_<unsigned long long>(epoch);
    p.n = n; p.P = (int)bufs.size(); p.rank = rank; p.scale = 1.0f / (float)p.P;
    p.fault = st ? &P_<OktState>(st)->fault : nullptr;
    p.timeout_ns = (unsigned long long)(timeout_s * 1e9);
    p.mc = P_<float>(mc);
    p.host_fault = P_<int>(host_fault);
    ck(launch_dense_allreduce(p, grid, S_(stream)), "dense allreduce launch");
}

static void kth_abs(uint64_t x, int n, int k, uint64_t st, uint64_t out, int grid, uint64_t stream) {
    ck(launch_kth_abs(P_<float>(x), n, k, P_<OktState>(st), P_<float>(out), grid, S_(stream)), "kth_abs launch");
}

static void fused_sgd(uint64_t p, uint64_t g, uint64_t mom, int n, double lr, double momentum, double dampening,
                      double wd, int nesterov, int first, int zero_grad, double grad_scale, uint64_t stream,
                      uint64_t lr_ptr, uint64_t fault_ptr) {
    ck(launch_fused_sgd(P_<float>(p), P_<float>(g), P_<float>(mom), n, (float)lr, (float)momentum, (float)dampening,
                        (float)wd, nesterov, first, zero_grad, (float)grad_scale, P_<float>(lr_ptr), P_<int>(fault_ptr),
                        S_(stream)),
       "fused_sgd");
}
static void fused_bert_adam(uint64_t p, uint64_t g, uint64_t m, uint64_t v, int n, double lr, double b1, double b2,
                            double eps, double wd, int zero_grad, uint64_t stream, uint64_t lr_ptr, uint64_t fault_ptr) {
    ck(launch_fused_bert_adam(P_<float>(p), P_<float>(g), P_<float>(m), P_<float>(v), n, (float)lr, (float)b1,
                              (float)b2, (float)eps, (float)wd, zero_grad, P_<float>(lr_ptr), P_<int>(fault_ptr),
                              S_(stream)),
       "fused_bert_adam");
}
static void bn_forward(uint64_t x, uint64_t y, uint64_t partial, uint64_t gamma, uint64_t beta, uint64_t cbias, uint64_t save_mean,
                       uint64_t save_invstd, uint64_t rmean, uint64_t rvar, uint64_t nbt, double momentum, double eps, int relu,
                       int M, int C, uint64_t stream) {
    if (C % 4 != 0) throw std::runtime_error("bn_forward: channel count must be a multiple of 4");
    ck(launch_bn_forward(P_<const float>(x), P_<float>(y), P_<float>(partial), P_<const float>(gamma), P_<const float>(beta),
                         P_<const float>(cbias), P_<float>(save_mean), P_<float>(save_invstd), P_<float>(rmean), P_<float>(rvar),
                         P_<long long>(nbt), (float)momentum, (float)eps, relu, M, C, S_(stream)), "bn_forward");
}
static void bn_backward(uint64_t x, uint64_t dy, uint64_t dx, uint64_t partial, uint64_t gamma, uint64_t beta, uint64_t save_mean,
                        uint64_t save_invstd, uint64_t dgamma, uint64_t dbeta, int relu, int M, int C, uint64_t stream) {
    ck(launch_bn_backward(P_<const float>(x), P_<const float>(dy), P_<float>(dx), P_<float>(partial), P_<const float>(gamma),
                          P_<const float>(beta), P_<const float>(save_mean), P_<const float>(save_invstd), P_<float>(dgamma),
                          P_<float>(dbeta), relu, M, C, S_(stream)), "bn_backward");
}
static void maxpool2_fwd(uint64_t x, uint64_t y, uint64_t arg, int N, int H, int W, int C, uint64_t stream) {
    if ((C % 4) || (H % 2) || (W % 2)) throw std::runtime_error("maxpool2_fwd: needs C % 4 == 0 and even H, W");
    ck(launch_maxpool2_fwd(P_<const float>(x), P_<float>(y), P_<unsigned char>(arg), N, H, W, C, S_(stream)), "maxpool2_fwd");
}
static void maxpool2_bwd(uint64_t dy, uint64_t arg, uint64_t dx, int N, int H, int W, int C, uint64_t stream) {
    ck(launch_maxpool2_bwd(P_<const float>(dy), P_<const unsigned char>(arg), P_<float>(dx), N, H, W, C, S_(stream)), "maxpool2_bwd");
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
          py::arg("grid"), py::arg("stream"), py::arg("st") = 0, py::arg("timeout_s") = 0.0, py::arg("mc") = 0,
          py::arg("host_fault") = 0);
    m.def("gtopk_run", &gtopk_run);
    m.def("land_grads", &land_grads);
    m.def("read_trace", &read_trace);
    m.def("host_flag_alloc", &host_flag_alloc);
    m.def("host_flag_free", &host_flag_free);
    m.def("host_flag_read", &host_flag_read);
    m.def("host_flag_clear", &host_flag_clear);
    m.def("fault_ptr", [](uint64_t st) { return (uint64_t)&P_<OktState>(st)->fault; });
    m.def("gather_max_coop_grid", &gather_max_coop_grid);
    m.def("gtopk_max_coop_grid", &gtopk_max_coop_grid);
    m.def("vmm_probe", &vmm_probe);
    m.def("vmm_create", &vmm_create);
    m.def("vmm_import", &vmm_import);
    m.def("vmm_map", &vmm_map);
    m.def("vmm_unmap", &vmm_unmap);
    m.def("vmm_release", &vmm_release);
    m.def("mc_create", &mc_create);
    m.def("mc_add_device", &mc_add_device);
    m.def("mc_bind", &mc_bind);
    m.def("close_fd", [](int fd) { if (fd >= 0) ::close(fd); });
    m.def("clear_fault", [](uint64_t st, uint64_t stream) {
        ck(cudaMemsetAsync(&P_<OktState>(st)->fault, 0, sizeof(int), S_(stream)), "clear_fault");
    });
    m.def("kth_abs", &kth_abs);
    m.def("fused_sgd", &fused_sgd, py::arg("p"), py::arg("g"), py::arg("mom"), py::arg("n"), py::arg("lr"),
          py::arg("momentum"), py::arg("dampening"), py::arg("wd"), py::arg("nesterov"), py::arg("first"),
          py::arg("zero_grad"), py::arg("grad_scale"), py::arg("stream"), py::arg("lr_ptr") = 0, py::arg("fault_ptr") = 0);
    m.def("fused_bert_adam", &fused_bert_adam, py::arg("p"), py::arg("g"), py::arg("m"), py::arg("v"), py::arg("n"),
          py::arg("lr"), py::arg("b1"), py::arg("b2"), py::arg("eps"), py::arg("wd"), py::arg("zero_grad"),
          py::arg("stream"), py::arg("lr_ptr") = 0, py::arg("fault_ptr") = 0);
    m.def("momentum_correct", &momentum_correct);
    m.def("maxpool2_fwd", &maxpool2_fwd);
    m.def("maxpool2_bwd", &maxpool2_bwd);
    m.def("bn_forward", &bn_forward);
    m.def("bn_backward", &bn_backward);
    m.def("bn_num_blocks", &bn_num_blocks);
    m.def("clip_by_norm", &clip_by_norm);
    m.attr("MAXP") = OKT_MAXP;
    m.attr("TRACE_LEN") = kTraceLen;
    m.attr("LAND_MAX") = kLandMax;
    m.attr("CHUNK") = kChunk;
}
