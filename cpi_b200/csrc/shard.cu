// Multi-GPU entry points of libcpi_b200.so (declared in include/cpi_b200.h): one process per GPU, window batches sharded
// contiguously over the ranks, ONE in-place NCCL all-gather of the fixed-size result records per batch (the
// reference builds one preintegrator per factor on one CPU thread, solvers/GraphSolver_IMU.cpp:43 -- windows share nothing).
//
// The preintegration kernel writes this rank's records straight into its slice of the caller's gather buffer; the
// all-gather runs on the communicator's OWN stream behind an event, so that the next batch's kernel (into a second gather
// buffer) overlaps it: the collective is ~0.3 ms for 8 x 10k fp64 records at NVLink rate and disappears behind a 0.9 ms kernel.
//
// NCCL is bound at run time (dlopen "libnccl.so.2": the copy torch already mapped when called from Python, the system one
// otherwise), so single-GPU users of the library need no NCCL at all.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstring>
#include <mutex>

#include "cpi_common.cuh"
#include "cpi_kernels.h"

extern "C" int cpi_preintegrate_batch(int model, int dtype, int64_t n_windows, const int64_t* sample_offsets, int64_t ns_uniform, const void* samples,
                                      const void* lin, const double* sigmas, int flags, void* out_records, void* stream);
extern "C" int cpi_record_doubles(int model);

namespace cpi { int capi_fail(int code, const char* fmt, ...); }

namespace {

struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

int nccl_load() {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.h) return CPI_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return cpi::capi_fail(CPI_ENODEVICE, "NCCL not found (dlopen libnccl.so.2): %s", dlerror());
    NcclApi a;
    a.h = h;
#define SYM(field, name) *(void**)(&a.field) = dlsym(h, name); if (!a.field) return cpi::capi_fail(CPI_ENODEVICE, "NCCL symbol %s missing", name)
    SYM(GetUniqueId, "ncclGetUniqueId"); SYM(CommInitRank, "ncclCommInitRank"); SYM(CommDestroy, "ncclCommDestroy");
    SYM(AllGather, "ncclAllGather"); SYM(GetErrorString, "ncclGetErrorString"); SYM(GetVersion, "ncclGetVersion");
#undef SYM
    g_nccl = a;
    return CPI_OK;
}
#define NC(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) return cpi::capi_fail(CPI_ECUDA, "%s failed: %s", #call, g_nccl.GetErrorString(r_)); } while (0)
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cpi::capi_fail(CPI_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

}  // namespace

struct cpi_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, device = 0;
    cudaStream_t stream = nullptr;            // communication stream
    cudaEvent_t kernel_done = nullptr;        // recorded on the caller's stream after the kernel
    static constexpr int NBUF = 4;
    void* buf[NBUF] = {nullptr};              // gather buffers seen recently ...
    cudaEvent_t gathered[NBUF] = {nullptr};   // ... and the event recorded behind their last all-gather
    int last = -1;
};

extern "C" {

int cpi_comm_unique_id(void* id_out) {
    if (!id_out) return cpi::capi_fail(CPI_EINVAL, "null pointer argument");
    int rc = nccl_load();
    if (rc) return rc;
    static_assert(sizeof(ncclUniqueId) == CPI_COMM_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    NC(g_nccl.GetUniqueId(&id));
    memcpy(id_out, &id, sizeof id);
    return CPI_OK;
}

int cpi_comm_create(const void* id_in, int rank, int world, cpi_comm** out) {
    if (!id_in || !out) return cpi::capi_fail(CPI_EINVAL, "null pointer argument");
    if (world < 1 || rank < 0 || rank >= world) return cpi::capi_fail(CPI_EINVAL, "bad rank %d / world %d", rank, world);
    int rc = nccl_load();
    if (rc) return rc;
    cpi_comm* c = new cpi_comm;
    c->rank = rank; c->world = world;
    CU(cudaGetDevice(&c->device));
    ncclUniqueId id;
    memcpy(&id, id_in, sizeof id);
    NC(g_nccl.CommInitRank(&c->comm, world, id, rank));
    // highest priority: when batch i's all-gather and batch i+1's kernel become runnable together, the collective's few CTAs must get
    // their SM slots first -- the preintegration kernel fills every SM in one wave and would otherwise starve it until it ends
    int prio_least = 0, prio_greatest = 0;
    CU(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    CU(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_greatest));
    CU(cudaEventCreateWithFlags(&c->kernel_done, cudaEventDisableTiming));
    for (int i = 0; i < cpi_comm::NBUF; i++) CU(cudaEventCreateWithFlags(&c->gathered[i], cudaEventDisableTiming));
    *out = c;
    return CPI_OK;
}

int cpi_comm_destroy(cpi_comm* c) {
    if (!c) return CPI_OK;
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->comm) g_nccl.CommDestroy(c->comm);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->kernel_done) cudaEventDestroy(c->kernel_done);
    for (int i = 0; i < cpi_comm::NBUF; i++) if (c->gathered[i]) cudaEventDestroy(c->gathered[i]);
    delete c;
    return CPI_OK;
}

int cpi_comm_rank(const cpi_comm* c) { return c ? c->rank : CPI_EINVAL; }
int cpi_comm_world(const cpi_comm* c) { return c ? c->world : CPI_EINVAL; }

int cpi_preintegrate_batch_sharded(cpi_comm* c, int model, int dtype, int64_t n_local, const int64_t* sample_offsets, int64_t ns_uniform,
                                   const void* samples, const void* lin, const double* sigmas, int flags, void* gather_records, void* stream) {
    if (!c || !gather_records) return cpi::capi_fail(CPI_EINVAL, "null pointer argument");
    if (n_local < 0) return cpi::capi_fail(CPI_EINVAL, "negative count");
    const int rd = cpi_record_doubles(model);
    if (rd < 0) return cpi::capi_fail(CPI_EINVAL, "model must be 1 or 2 (got %d)", model);
    if (dtype != 64 && dtype != 32) return cpi::capi_fail(CPI_EINVAL, "dtype must be 64 or 32 (got %d)", dtype);
    int dev = 0;
    CU(cudaGetDevice(&dev));
    if (dev != c->device) return cpi::capi_fail(CPI_EINVAL, "communicator was created on device %d, current device is %d", c->device, dev);
    const size_t slice = (size_t)n_local * rd * (dtype == 32 ? 4 : 8);
    cudaStream_t st = (cudaStream_t)stream;
    // this gather buffer may still be the source / destination of an earlier all-gather: order the kernel behind it
    int slot = -1;
    for (int i = 0; i < cpi_comm::NBUF; i++) if (c->buf[i] == gather_records) slot = i;
    if (slot >= 0) CU(cudaStreamWaitEvent(st, c->gathered[slot], 0));
    else { slot = (c->last + 1) % cpi_comm::NBUF; c->buf[slot] = gather_records; }
    int rc = cpi_preintegrate_batch(model, dtype, n_local, sample_offsets, ns_uniform, samples, lin, sigmas, flags,
                                    (char*)gather_records + (size_t)c->rank * slice, stream);
    if (rc) return rc;
    if (c->world > 1) {
        CU(cudaEventRecord(c->kernel_done, st));
        CU(cudaStreamWaitEvent(c->stream, c->kernel_done, 0));
        if (slice > 0)
            NC(g_nccl.AllGather((const char*)gather_records + (size_t)c->rank * slice, gather_records, slice, ncclChar, c->comm, c->stream));
        CU(cudaEventRecord(c->gathered[slot], c->stream));
    } else {
        CU(cudaEventRecord(c->gathered[slot], st));
    }
    c->last = slot;
    return CPI_OK;
}

int cpi_comm_wait(cpi_comm* c, void* stream) {
    if (!c) return cpi::capi_fail(CPI_EINVAL, "null pointer argument");
    if (c->last >= 0) CU(cudaStreamWaitEvent((cudaStream_t)stream, c->gathered[c->last], 0));
    return CPI_OK;
}

}  // extern "C"
