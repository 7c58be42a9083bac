// nccl_halo.cu -- the slab halo exchange driven from C: one ncclGroupStart / ncclSend / ncclRecv /
// ncclGroupEnd batch per exchange phase, enqueued on the caller's stream.
//
// Why not torch.distributed's P2P ops: one batch_isend_irecv of the u, v, density halos costs ~130 us of host time
// (~25 us per P2POp; tools/slab_host_time.py), three exchanges per step ~250 us -- more than the ~260 us of kernels
// a rank of an 8-slab 8192^2 grid runs per step, so the multi-GPU step was bound by the Python/ProcessGroup host
// path, not by NVLink or the GPU.  The same sends and receives issued here cost one C call per phase.
//
// NCCL is not linked: the library that is already in the process (PyTorch's bundled libnccl.so.2) is looked up with
// dlopen, so libsmoke_sm100.so keeps loading on machines without NCCL, where these entry points return
// SMK_EUNSUPPORTED.  The communicator is the library's own (ncclCommInitRank with an id that the caller broadcasts
// over its existing torch.distributed group); ownership of every buffer stays with the caller.
#include <dlfcn.h>
#include <cstring>
#include "common.cuh"

namespace smk {

struct NcclUniqueId { char internal[128]; };            // nccl.h: NCCL_UNIQUE_ID_BYTES
typedef void* NcclComm;
enum { kNcclFloat32 = 7 };                               // nccl.h: ncclFloat32

struct NcclApi {
    void* handle = nullptr;
    int (*GetVersion)(int*) = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static int nccl_fail(const char* what, int rc)
{
    return fail(1000 + rc, "%s: NCCL error %d (%s)", what, rc, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
}

}  // namespace smk

using namespace smk;

extern "C" {

int smk_nccl_load(const char* libpath_host, int32_t* version_host)
{
    if (!g_nccl.handle) {
        void* h = nullptr;
        if (libpath_host && libpath_host[0]) h = dlopen(libpath_host, RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);      // the copy PyTorch already brought into the process
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return fail(SMK_EUNSUPPORTED, "smk_nccl_load: libnccl.so.2 not found (%s)", dlerror());
        NcclApi api;
        api.handle = h;
#define SMK_SYM(field, name) *reinterpret_cast<void**>(&api.field) = dlsym(h, name); \
        if (!api.field) return fail(SMK_EUNSUPPORTED, "smk_nccl_load: symbol %s missing", name);
        SMK_SYM(GetVersion, "ncclGetVersion") SMK_SYM(GetUniqueId, "ncclGetUniqueId") SMK_SYM(CommInitRank, "ncclCommInitRank")
        SMK_SYM(CommDestroy, "ncclCommDestroy") SMK_SYM(GroupStart, "ncclGroupStart") SMK_SYM(GroupEnd, "ncclGroupEnd")
        SMK_SYM(Send, "ncclSend") SMK_SYM(Recv, "ncclRecv") SMK_SYM(GetErrorString, "ncclGetErrorString")
#undef SMK_SYM
        g_nccl = api;
    }
    if (version_host) {
        int v = 0;
        const int rc = g_nccl.GetVersion(&v);
        if (rc) return nccl_fail("ncclGetVersion", rc);
        *version_host = v;
    }
    return SMK_OK;
}

int smk_nccl_unique_id(void* id128_host)
{
    if (!g_nccl.handle) return fail(SMK_EUNSUPPORTED, "smk_nccl_unique_id: call smk_nccl_load first");
    if (!id128_host) return fail(SMK_EINVAL, "smk_nccl_unique_id: NULL");
    NcclUniqueId id;
    const int rc = g_nccl.GetUniqueId(&id);
    if (rc) return nccl_fail("ncclGetUniqueId", rc);
    memcpy(id128_host, &id, sizeof(id));
    return SMK_OK;
}

int smk_nccl_comm_init(const void* id128_host, int32_t rank, int32_t world, void** comm_out_host)
{
    if (!g_nccl.handle) return fail(SMK_EUNSUPPORTED, "smk_nccl_comm_init: call smk_nccl_load first");
    if (!id128_host || !comm_out_host || world < 1 || rank < 0 || rank >= world) return fail(SMK_EINVAL, "smk_nccl_comm_init: bad arguments");
    NcclUniqueId id;
    memcpy(&id, id128_host, sizeof(id));
    NcclComm comm = nullptr;
    const int rc = g_nccl.CommInitRank(&comm, world, id, rank);
    if (rc) return nccl_fail("ncclCommInitRank", rc);
    *comm_out_host = comm;
    return SMK_OK;
}

int smk_nccl_comm_destroy(void* comm)
{
    if (!g_nccl.handle || !comm) return SMK_OK;
    const int rc = g_nccl.CommDestroy(comm);
    return rc ? nccl_fail("ncclCommDestroy", rc) : SMK_OK;
}

int smk_nccl_exchange(void* comm, const smk_halo_block_t* blocks_host, int32_t nblocks, void* stream)
{
    if (!g_nccl.handle) return fail(SMK_EUNSUPPORTED, "smk_nccl_exchange: call smk_nccl_load first");
    if (!comm || (nblocks > 0 && !blocks_host) || nblocks < 0) return fail(SMK_EINVAL, "smk_nccl_exchange: bad arguments");
    for (int k = 0; k < nblocks; ++k)
        if (!blocks_host[k].ptr || blocks_host[k].count < 0 || blocks_host[k].peer < 0)
            return fail(SMK_EINVAL, "smk_nccl_exchange: block %d is malformed", k);
    if (nblocks == 0) return SMK_OK;
    int rc = g_nccl.GroupStart();
    if (rc) return nccl_fail("ncclGroupStart", rc);
    for (int k = 0; k < nblocks && !rc; ++k) {
        const smk_halo_block_t& b = blocks_host[k];
        if (b.count == 0) continue;
        rc = b.is_send ? g_nccl.Send(b.ptr, (size_t)b.count, kNcclFloat32, b.peer, comm, (cudaStream_t)stream)
                       : g_nccl.Recv(b.ptr, (size_t)b.count, kNcclFloat32, b.peer, comm, (cudaStream_t)stream);
    }
    const int rc_end = g_nccl.GroupEnd();
    if (rc) return nccl_fail("ncclSend/ncclRecv", rc);
    if (rc_end) return nccl_fail("ncclGroupEnd", rc_end);
    return SMK_OK;
}

}  // extern "C"
