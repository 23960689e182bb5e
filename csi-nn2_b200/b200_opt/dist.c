/*
 * dist.c -- the multi-GPU plumbing of the b200 backend, host side in C.
 *
 * Images are independent units, so N GPUs = N processes x one session each on its own shard of the batch
 * and NO collective on the inference path.  The one exchange is at start-up: rank `root` packs the weights
 * (quantisation tables, packed kernels, border tables -- the whole weight arena of the session, built in the
 * same order on every rank), the other ranks allocate the arena without uploading
 * (SHL_B200_SKIP_WEIGHT_UPLOAD) and receive it with ONE ncclBroadcast over NVLink / NVSwitch.
 *
 * NCCL is bound at run time (dlopen of libnccl.so.2 -- the copy the host process already loaded, e.g. the
 * one bundled with PyTorch, is reused), so libshl_b200.so has no link-time dependency on it and single-GPU
 * users never load it.  The communicator is this library's own: ncclGetUniqueId on one rank, the 128-byte
 * id carried to the others by whatever the host already has (MPI, torch.distributed, a file), then
 * ncclCommInitRank on every rank.
 */
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include "b200_internal.h"

typedef struct ncclComm *ncclComm_t;
typedef struct {
    char internal[128];
} ncclUniqueId;
typedef int ncclResult_t; /* ncclSuccess == 0 */
enum { k_ncclUint8 = 1 };  /* ncclDataType_t: ncclInt8 = 0, ncclUint8 = 1 (nccl.h) */

static struct {
    void *lib;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, void *);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    const char *(*GetErrorString)(ncclResult_t);
} g_nccl;

static int nccl_load(void)
{
    if (g_nccl.lib) return CSINN_TRUE;
    const char *names[] = {"libnccl.so.2", "libnccl.so", NULL};
    for (int i = 0; names[i] && !g_nccl.lib; i++) g_nccl.lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!g_nccl.lib) {
        b200_fail("NCCL not found (dlopen libnccl.so.2: %s)", dlerror());
        return CSINN_FALSE;
    }
    *(void **)&g_nccl.GetUniqueId = dlsym(g_nccl.lib, "ncclGetUniqueId");
    *(void **)&g_nccl.CommInitRank = dlsym(g_nccl.lib, "ncclCommInitRank");
    *(void **)&g_nccl.Broadcast = dlsym(g_nccl.lib, "ncclBroadcast");
    *(void **)&g_nccl.CommDestroy = dlsym(g_nccl.lib, "ncclCommDestroy");
    *(void **)&g_nccl.GetErrorString = dlsym(g_nccl.lib, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.Broadcast || !g_nccl.CommDestroy) {
        b200_fail("libnccl lacks ncclGetUniqueId / ncclCommInitRank / ncclBroadcast / ncclCommDestroy");
        dlclose(g_nccl.lib);
        memset(&g_nccl, 0, sizeof(g_nccl));
        return CSINN_FALSE;
    }
    return CSINN_TRUE;
}

static int nccl_ok(ncclResult_t r, const char *what)
{
    if (r == 0) return CSINN_TRUE;
    b200_fail("%s -> NCCL error %d (%s)", what, r, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    return CSINN_FALSE;
}

int shl_b200_nccl_unique_id(void *id128)
{
    if (!id128 || nccl_load() != CSINN_TRUE) return CSINN_FALSE;
    return nccl_ok(g_nccl.GetUniqueId((ncclUniqueId *)id128), "ncclGetUniqueId");
}

/* the communicator of this rank, bound to the session's GPU */
int shl_b200_nccl_comm_init(struct csinn_session *sess, const void *id128, int rank, int world, void **comm)
{
    b200_option *opt = b200_option_of(sess);
    if (!opt || !id128 || !comm || nccl_load() != CSINN_TRUE) {
        if (!opt) b200_fail("nccl_comm_init: not a b200 graph session");
        return CSINN_FALSE;
    }
    b200_set_device(opt->ctx.device);
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t c = NULL;
    if (nccl_ok(g_nccl.CommInitRank(&c, world, id, rank), "ncclCommInitRank") != CSINN_TRUE) return CSINN_FALSE;
    *comm = c;
    return CSINN_TRUE;
}

/* ONE collective for the whole run: the weight arena of `root` into every rank's arena (same size everywhere:
 * the arenas are bump-allocated in graph order).  Synchronous: returns when this rank's copy is complete. */
int shl_b200_session_broadcast_weights(struct csinn_session *sess, void *comm, int root)
{
    b200_option *opt = b200_option_of(sess);
    if (!opt || !opt->ctx.wbase || !comm || nccl_load() != CSINN_TRUE) {
        b200_fail("broadcast_weights: session not set up, or no communicator");
        return CSINN_FALSE;
    }
    b200_set_device(opt->ctx.device);
    if (nccl_ok(g_nccl.Broadcast(opt->ctx.wbase, opt->ctx.wbase, opt->ctx.wused, k_ncclUint8, root, (ncclComm_t)comm,
                                 opt->ctx.stream),
                "ncclBroadcast(weight arena)") != CSINN_TRUE)
        return CSINN_FALSE;
    if (b200_stream_sync(opt->ctx.stream) != B200_OK) {
        b200_fail("broadcast_weights: %s", b200_last_error());
        return CSINN_FALSE;
    }
    opt->ctx.skip_upload = 0;
    return CSINN_TRUE;
}

int shl_b200_nccl_comm_destroy(void *comm)
{
    if (!comm || !g_nccl.lib) return CSINN_TRUE;
    return nccl_ok(g_nccl.CommDestroy((ncclComm_t)comm), "ncclCommDestroy");
}
