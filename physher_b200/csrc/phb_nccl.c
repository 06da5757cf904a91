/*
 * phb_nccl.c -- the one collective of the pattern-sharded tree likelihood, in the C host layer.
 *
 * Site patterns shard across GPUs; the only exchange is ONE ncclAllReduce(sum, double) of [lnL, grad[N], inf flag] per
 * evaluation, enqueued on the stream the evaluation ran on -- no host synchronisation between the kernels and the collective,
 * no host-side sum.  Two hosts use it:
 *   - one process per GPU (torchtree-physher, MPI-style launchers, bench.py under torchrun): phb_comm_unique_id on rank 0, the
 *     128 bytes travel by whatever the launcher offers, phb_comm_init_rank on every rank, phb_tlk_gradient_allreduce[_device];
 *   - one process, all GPUs (physher itself, src/physher.c:61-322 is a single-threaded C program): phb_group (phb_group.c) builds
 *     its communicators with ncclCommInitAll and issues the per-shard all-reduces inside one ncclGroupStart / ncclGroupEnd.
 *
 * NCCL is bound at run time (dlopen "libnccl.so.2": the system library, or the copy a host process such as PyTorch has already
 * loaded under the same soname), so the library itself loads on machines without NCCL; asking for a communicator there fails
 * loudly with PHB_ESTATE.
 */
#include "../../include/physher_b200.h"

#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

int phb_internal_fail(int code, const char *msg); /* phb_treelikelihood.c: sets phb_last_error() */

/* the slice of nccl.h this file needs (layout-compatible restatement: ncclUniqueId is 128 opaque bytes passed by value) */
typedef struct { char internal[PHB_NCCL_ID_BYTES]; } nccl_unique_id;
typedef void *nccl_comm;
enum { NCCL_SUM = 0, NCCL_DOUBLE = 8 };

static struct {
	void *handle;
	int tried;
	int (*GetVersion)(int *);
	int (*GetUniqueId)(nccl_unique_id *);
	int (*CommInitRank)(nccl_comm *, int, nccl_unique_id, int);
	int (*CommInitAll)(nccl_comm *, int, const int *);
	int (*CommDestroy)(nccl_comm);
	int (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm, cudaStream_t);
	int (*GroupStart)(void);
	int (*GroupEnd)(void);
	const char *(*GetErrorString)(int);
} nccl;

static int nccl_load(void) {
	if (nccl.tried) return nccl.handle ? PHB_OK : PHB_ESTATE;
	nccl.tried = 1;
	const char *names[] = {"libnccl.so.2", "libnccl.so"};
	void *h = NULL;
	for (int i = 0; i < 2 && !h; i++) h = dlopen(names[i], RTLD_NOW | RTLD_LOCAL);
	if (!h) return phb_internal_fail(PHB_ESTATE, "NCCL is not loadable (libnccl.so.2): multi-GPU collectives are unavailable");
#define BIND(field, sym)                                                                   \
	do {                                                                                   \
		*(void **)(&nccl.field) = dlsym(h, sym);                                           \
		if (!nccl.field) {                                                                 \
			dlclose(h);                                                                    \
			return phb_internal_fail(PHB_ESTATE, "libnccl.so.2 does not export " sym);     \
		}                                                                                  \
	} while (0)
	BIND(GetVersion, "ncclGetVersion");
	BIND(GetUniqueId, "ncclGetUniqueId");
	BIND(CommInitRank, "ncclCommInitRank");
	BIND(CommInitAll, "ncclCommInitAll");
	BIND(CommDestroy, "ncclCommDestroy");
	BIND(AllReduce, "ncclAllReduce");
	BIND(GroupStart, "ncclGroupStart");
	BIND(GroupEnd, "ncclGroupEnd");
	BIND(GetErrorString, "ncclGetErrorString");
#undef BIND
	nccl.handle = h;
	return PHB_OK;
}

static int nccl_fail(const char *what, int rc) {
	char msg[256];
	snprintf(msg, sizeof(msg), "%s: %s", what, nccl.GetErrorString ? nccl.GetErrorString(rc) : "NCCL error");
	return phb_internal_fail(PHB_ECUDA, msg);
}

int phb_nccl_version(void) {
	int v = 0;
	if (nccl_load() != PHB_OK || nccl.GetVersion(&v) != 0) return 0;
	return v;
}

struct phb_comm {
	nccl_comm comm;
	int rank, nranks, device;
};

int phb_comm_unique_id(void *id) {
	if (!id) return phb_internal_fail(PHB_EINVAL, "phb_comm_unique_id: id is required");
	int rc = nccl_load();
	if (rc) return rc;
	nccl_unique_id u;
	if ((rc = nccl.GetUniqueId(&u))) return nccl_fail("ncclGetUniqueId", rc);
	memcpy(id, u.internal, PHB_NCCL_ID_BYTES);
	return PHB_OK;
}

phb_comm *phb_comm_init_rank(int nranks, int rank, const void *id, int device) {
	if (nranks < 1 || rank < 0 || rank >= nranks || !id) {
		phb_internal_fail(PHB_EINVAL, "phb_comm_init_rank: need 0 <= rank < nranks and the unique id of rank 0");
		return NULL;
	}
	if (nccl_load()) return NULL;
	if (cudaSetDevice(device) != cudaSuccess) {
		phb_internal_fail(PHB_ECUDA, "phb_comm_init_rank: cudaSetDevice failed");
		return NULL;
	}
	phb_comm *c = (phb_comm *)calloc(1, sizeof(phb_comm));
	if (!c) return NULL;
	nccl_unique_id u;
	memcpy(u.internal, id, PHB_NCCL_ID_BYTES);
	const int rc = nccl.CommInitRank(&c->comm, nranks, u, rank);
	if (rc) {
		nccl_fail("ncclCommInitRank", rc);
		free(c);
		return NULL;
	}
	c->rank = rank, c->nranks = nranks, c->device = device;
	return c;
}

void phb_comm_free(phb_comm *c) {
	if (!c) return;
	if (c->comm && nccl.CommDestroy) {
		cudaSetDevice(c->device);
		nccl.CommDestroy(c->comm);
	}
	free(c);
}

int phb_comm_size(const phb_comm *c) { return c ? c->nranks : 0; }
int phb_comm_rank(const phb_comm *c) { return c ? c->rank : -1; }

/* in-place SUM all-reduce of `count` doubles on `stream` (enqueue only) */
int phb_internal_allreduce(phb_comm *c, double *buf, size_t count, void *stream) {
	if (!c || !c->comm) return phb_internal_fail(PHB_EINVAL, "all-reduce without a communicator");
	if (cudaSetDevice(c->device) != cudaSuccess) return phb_internal_fail(PHB_ECUDA, "cudaSetDevice failed");
	const int rc = nccl.AllReduce(buf, buf, count, NCCL_DOUBLE, NCCL_SUM, c->comm, (cudaStream_t)stream);
	if (rc) return nccl_fail("ncclAllReduce", rc);
	return PHB_OK;
}

/* ---- single process, several devices: the communicators of a phb_group ------------------------------------------------------ */

/* one communicator per device of `devices` (all distinct); comms[] receives handles that phb_internal_comms_free releases */
int phb_internal_comms_init_all(int n, const int *devices, phb_comm **comms) {
	int rc = nccl_load();
	if (rc) return rc;
	nccl_comm *raw = (nccl_comm *)calloc(n, sizeof(nccl_comm));
	if (!raw) return phb_internal_fail(PHB_ENOMEM, "out of memory");
	if ((rc = nccl.CommInitAll(raw, n, devices))) {
		free(raw);
		return nccl_fail("ncclCommInitAll", rc);
	}
	for (int i = 0; i < n; i++) {
		comms[i] = (phb_comm *)calloc(1, sizeof(phb_comm));
		if (!comms[i]) {
			free(raw);
			return phb_internal_fail(PHB_ENOMEM, "out of memory");
		}
		comms[i]->comm = raw[i], comms[i]->rank = i, comms[i]->nranks = n, comms[i]->device = devices[i];
	}
	free(raw);
	return PHB_OK;
}

/* the per-shard all-reduces of one evaluation as ONE NCCL group (a single thread drives every rank) */
int phb_internal_allreduce_all(int n, phb_comm **comms, double **bufs, size_t count, void **streams) {
	int rc = nccl.GroupStart();
	if (rc) return nccl_fail("ncclGroupStart", rc);
	for (int i = 0; i < n; i++) {
		rc = nccl.AllReduce(bufs[i], bufs[i], count, NCCL_DOUBLE, NCCL_SUM, comms[i]->comm, (cudaStream_t)streams[i]);
		if (rc) {
			nccl.GroupEnd();
			return nccl_fail("ncclAllReduce", rc);
		}
	}
	if ((rc = nccl.GroupEnd())) return nccl_fail("ncclGroupEnd", rc);
	return PHB_OK;
}
