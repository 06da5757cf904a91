// phb_patterns.cu -- site-pattern compression on the device.
//
// Replaces new_SitePattern2 / _make_patterns (sitepattern.c:186-251, 731-754): the unique columns of an encoded
// alignment, their multiplicities, and -- so that a drop-in sees the SAME pattern order as the reference and per-pattern
// arrays (sp->weights, tlk->pattern_lk) line up index by index -- the order in which the reference's chained hash table
// hands them back (Hashtable_next, hashtable.c:414-451).  Integer work: bit-exact.
//
//   device  one thread per site: the reference's own column hash (hashtable_hash_uint8_t, sitepattern.c:71-79, through
//           the scrambler hashfn, hashtable.c:188-197) plus an independent hash;
//           stable radix sort of the sites by the 64-bit key (reference hash | 32 independent bits) (cub), run heads confirmed by comparing whole columns
//           (a key collision between different columns is detected, never silently merged);
//           runs ordered by their first site = the order in which the reference inserts new keys;
//   host    replay of the reference's table policy on the per-pattern hash values only (insert at the bucket head,
//           grow through its prime list at load factor 0.65 with its list-reversing rehash, iterate buckets in index
//           order): O(P) integer work, no column data;
//   device  gather of the representative columns and counts in that order, and the site -> pattern map.
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/physher_b200.h"

static thread_local char pat_err[256];
extern "C" const char *phb_patterns_last_error(void) { return pat_err; }

#define PAT_CHECK(call)                                                                                  \
	do {                                                                                                 \
		cudaError_t e__ = (call);                                                                        \
		if (e__ != cudaSuccess) {                                                                        \
			snprintf(pat_err, sizeof(pat_err), "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
			rc = PHB_ECUDA;                                                                              \
			goto done;                                                                                   \
		}                                                                                                \
	} while (0)

// aln: [T][nsites] encoded states, sequence order as in the alignment (the hash runs over taxa in that order)
__global__ void k_pat_hash(int T, size_t nsites, const uint8_t *__restrict__ aln, uint64_t *__restrict__ key, uint32_t *__restrict__ site) {
	const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= nsites) return;
	uint32_t h = aln[s];                 // hashtable_hash_uint8_t: hash = values[0]
	uint64_t g = 0xcbf29ce484222325ull;  // independent FNV-1a style hash with a 64-bit finaliser
	g = (g ^ aln[s]) * 0x100000001b3ull;
	for (int t = 1; t < T; t++) {
		const uint32_t v = aln[(size_t)t * nsites + s];
		h ^= v + 0x9e3779b9u + (h << 6) + (h >> 2);
		g = (g ^ v) * 0x100000001b3ull;
		g ^= g >> 29;
	}
	// hashfn (hashtable.c:188-197)
	h += ~(h << 9);
	h ^= ((h >> 14) | (h << 18));
	h += (h << 4);
	h ^= ((h >> 10) | (h << 22));
	g ^= g >> 32;
	g *= 0xd6e8feb86659fd93ull;
	g ^= g >> 32;
	key[s] = ((uint64_t)h << 32) | (g >> 32);  // reference hash in the high word, 32 independent bits below
	site[s] = (uint32_t)s;
}

// sorted order: head[i] = 1 when element i opens a new run; equal keys are confirmed on the column data
__global__ void k_pat_heads(int T, size_t nsites, const uint8_t *__restrict__ aln, const uint64_t *__restrict__ key,
                            const uint32_t *__restrict__ site, uint32_t *__restrict__ head, int *__restrict__ collision) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nsites) return;
	uint32_t hd = 1;
	if (i > 0 && key[i] == key[i - 1]) {
		const size_t a = site[i], b = site[i - 1];
		bool same = true;
		for (int t = 0; t < T && same; t++) same = aln[(size_t)t * nsites + a] == aln[(size_t)t * nsites + b];
		if (same) hd = 0;
		else *collision = 1;  // two different columns share a 64-bit key: refuse rather than risk a split run
	}
	head[i] = hd;
}

// run r (in key order): first site (= its smallest site, the sort is stable), its key hash, its length
__global__ void k_pat_runs(size_t nsites, const uint32_t *__restrict__ head, const uint32_t *__restrict__ runid /* inclusive scan of head */,
                           const uint32_t *__restrict__ site, const uint64_t *__restrict__ key, uint32_t *__restrict__ run_first,
                           uint32_t *__restrict__ run_hash, uint32_t *__restrict__ run_start, uint32_t *__restrict__ run_index) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nsites) return;
	if (head[i]) {
		const uint32_t r = runid[i] - 1;
		run_first[r] = site[i];
		run_hash[r] = (uint32_t)(key[i] >> 32);
		run_start[r] = (uint32_t)i;
		run_index[r] = r;
	}
}

// patterns[t][k] and weights[k] for reference position k; pos_of_run[r] = k
__global__ void k_pat_gather(int T, size_t nsites, uint32_t npat, const uint8_t *__restrict__ aln, const uint32_t *__restrict__ order /* [k] -> run */,
                             const uint32_t *__restrict__ run_first, const uint32_t *__restrict__ run_start, uint8_t *__restrict__ patterns,
                             double *__restrict__ weights, uint32_t *__restrict__ pos_of_run) {
	const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= npat) return;
	const uint32_t r = order[k];
	const size_t s = run_first[r];
	for (int t = 0; t < T; t++) patterns[(size_t)t * npat + k] = aln[(size_t)t * nsites + s];
	const uint32_t end = r + 1 < npat ? run_start[r + 1] : (uint32_t)nsites;
	weights[k] = (double)(end - run_start[r]);
	pos_of_run[r] = (uint32_t)k;
}

__global__ void k_pat_site_map(size_t nsites, const uint32_t *__restrict__ runid, const uint32_t *__restrict__ site,
                               const uint32_t *__restrict__ pos_of_run, int *__restrict__ site_to_pattern) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nsites) return;
	site_to_pattern[site[i]] = (int)pos_of_run[runid[i] - 1];
}

// Replay of the reference's hash table on hash values only: keys arrive in first-appearance order, new keys go to the head
// of bucket hash % size (Hashtable_add, hashtable.c:262-312), the table grows when length == ceil(size * 0.65) BEFORE the
// insertion, moving entries bucket by bucket from the old heads to the new heads (Hashtable_expand :199-249), and the
// iterator walks the buckets in index order, each chain from its head (Hashtable_next :426-451).
static int replay_reference_table(uint32_t npat, const uint32_t *hash, unsigned initial_size, uint32_t *order) {
	static const unsigned primes[] = {5,        53,       97,       193,      389,       769,       1543,      3079,      6151,
	                                  12289,    24593,    49157,    98317,    196613,    393241,    786433,    1572869,   3145739,
	                                  6291469,  12582917, 25165843, 50331653, 100663319, 201326611, 402653189, 805306457, 1610612741};
	const int nprimes = (int)(sizeof(primes) / sizeof(primes[0]));
	int pi = 0;
	while (pi < nprimes - 1 && primes[pi] < initial_size) pi++;
	unsigned size = primes[pi];
	unsigned loadlimit = (unsigned)ceil(size * 0.65);
	int32_t *table = (int32_t *)malloc(sizeof(int32_t) * size);
	int32_t *next = (int32_t *)malloc(sizeof(int32_t) * (npat ? npat : 1));
	if (!table || !next) {
		free(table), free(next);
		return PHB_ENOMEM;
	}
	for (unsigned i = 0; i < size; i++) table[i] = -1;
	for (uint32_t q = 0; q < npat; q++) {
		if (q == loadlimit && pi + 1 < nprimes) {  // hash->length == hash->loadlimit
			const unsigned newsize = primes[++pi];
			int32_t *nt = (int32_t *)malloc(sizeof(int32_t) * newsize);
			if (!nt) {
				free(table), free(next);
				return PHB_ENOMEM;
			}
			for (unsigned i = 0; i < newsize; i++) nt[i] = -1;
			for (unsigned i = 0; i < size; i++) {
				int32_t e;
				while ((e = table[i]) >= 0) {
					table[i] = next[e];
					const unsigned idx = hash[e] % newsize;
					next[e] = nt[idx];
					nt[idx] = e;
				}
			}
			free(table);
			table = nt;
			size = newsize;
			loadlimit = (unsigned)ceil(size * 0.65);
		}
		const unsigned idx = hash[q] % size;
		next[q] = table[idx];
		table[idx] = (int32_t)q;
	}
	uint32_t k = 0;
	for (unsigned i = 0; i < size; i++)
		for (int32_t e = table[i]; e >= 0; e = next[e]) order[k++] = (uint32_t)e;
	free(table);
	free(next);
	return k == npat ? PHB_OK : PHB_EINVAL;
}

extern "C" int phb_compress_patterns(int device, int ntaxa, size_t nsites, const uint8_t *alignment, int hashtable_size, size_t *npatterns,
                                     uint8_t **patterns, double **weights, int **site_to_pattern) {
	int rc = PHB_OK;
	pat_err[0] = 0;
	if (ntaxa < 1 || nsites < 1 || nsites >= 0x7fffffffull || !alignment || !npatterns || !patterns || !weights) {
		snprintf(pat_err, sizeof(pat_err), "phb_compress_patterns: bad arguments");
		return PHB_EINVAL;
	}
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
		snprintf(pat_err, sizeof(pat_err), "no CUDA device visible: the B200 path has no CPU fallback");
		return PHB_ECUDA;
	}
	const size_t n = nsites;
	const unsigned blocks = (unsigned)((n + 255) / 256);
	uint8_t *d_aln = NULL, *d_pat = NULL;
	uint64_t *d_key = NULL, *d_key2 = NULL;
	uint32_t *d_site = NULL, *d_site2 = NULL, *d_head = NULL, *d_runid = NULL;
	uint32_t *d_run_first = NULL, *d_run_hash = NULL, *d_run_start = NULL, *d_run_index = NULL, *d_first_sorted = NULL, *d_index_sorted = NULL;
	uint32_t *d_order = NULL, *d_pos = NULL;
	double *d_w = NULL;
	int *d_map = NULL, *d_coll = NULL;
	void *d_tmp = NULL;
	size_t tmp_bytes = 0, need = 0;
	uint32_t *h_hash = NULL, *h_index = NULL, *h_order = NULL, *h_hash_q = NULL;
	uint32_t npat = 0;
	int coll = 0;
	cudaStream_t st = 0;
	*patterns = NULL, *weights = NULL;
	if (site_to_pattern) *site_to_pattern = NULL;

	PAT_CHECK(cudaSetDevice(device));
	PAT_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
	PAT_CHECK(cudaMalloc((void **)&d_aln, (size_t)ntaxa * n));
	PAT_CHECK(cudaMalloc((void **)&d_key, n * 8));
	PAT_CHECK(cudaMalloc((void **)&d_key2, n * 8));
	PAT_CHECK(cudaMalloc((void **)&d_site, n * 4));
	PAT_CHECK(cudaMalloc((void **)&d_site2, n * 4));
	PAT_CHECK(cudaMalloc((void **)&d_head, n * 4));
	PAT_CHECK(cudaMalloc((void **)&d_runid, n * 4));
	PAT_CHECK(cudaMalloc((void **)&d_coll, sizeof(int)));
	PAT_CHECK(cudaMemcpyAsync(d_aln, alignment, (size_t)ntaxa * n, cudaMemcpyHostToDevice, st));
	PAT_CHECK(cudaMemsetAsync(d_coll, 0, sizeof(int), st));
	k_pat_hash<<<blocks, 256, 0, st>>>(ntaxa, n, d_aln, d_key, d_site);

	PAT_CHECK(cub::DeviceRadixSort::SortPairs(NULL, need, d_key, d_key2, d_site, d_site2, (int)n, 0, 64, st));
	tmp_bytes = need;
	PAT_CHECK(cub::DeviceScan::InclusiveSum(NULL, need, d_head, d_runid, (int)n, st));
	if (need > tmp_bytes) tmp_bytes = need;
	PAT_CHECK(cudaMalloc(&d_tmp, tmp_bytes));
	need = tmp_bytes;
	// stable sort: equal keys keep ascending site order, so a run's first element is the column's first occurrence
	PAT_CHECK(cub::DeviceRadixSort::SortPairs(d_tmp, need, d_key, d_key2, d_site, d_site2, (int)n, 0, 64, st));
	k_pat_heads<<<blocks, 256, 0, st>>>(ntaxa, n, d_aln, d_key2, d_site2, d_head, d_coll);
	need = tmp_bytes;
	PAT_CHECK(cub::DeviceScan::InclusiveSum(d_tmp, need, d_head, d_runid, (int)n, st));
	PAT_CHECK(cudaMemcpyAsync(&npat, d_runid + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
	PAT_CHECK(cudaMemcpyAsync(&coll, d_coll, sizeof(int), cudaMemcpyDeviceToHost, st));
	PAT_CHECK(cudaStreamSynchronize(st));
	if (coll) {
		snprintf(pat_err, sizeof(pat_err), "phb_compress_patterns: two different columns share a 64-bit key (refusing to merge)");
		rc = PHB_ESTATE;
		goto done;
	}
	PAT_CHECK(cudaMalloc((void **)&d_run_first, (size_t)npat * 4));
	PAT_CHECK(cudaMalloc((void **)&d_run_hash, (size_t)npat * 4));
	PAT_CHECK(cudaMalloc((void **)&d_run_start, (size_t)npat * 4));
	PAT_CHECK(cudaMalloc((void **)&d_run_index, (size_t)npat * 4));
	PAT_CHECK(cudaMalloc((void **)&d_first_sorted, (size_t)npat * 4));
	PAT_CHECK(cudaMalloc((void **)&d_index_sorted, (size_t)npat * 4));
	PAT_CHECK(cudaMalloc((void **)&d_order, (size_t)npat * 4));
	PAT_CHECK(cudaMalloc((void **)&d_pos, (size_t)npat * 4));
	k_pat_runs<<<blocks, 256, 0, st>>>(n, d_head, d_runid, d_site2, d_key2, d_run_first, d_run_hash, d_run_start, d_run_index);
	// runs in order of first appearance = the order in which the reference inserts new keys
	need = 0;
	PAT_CHECK(cub::DeviceRadixSort::SortPairs(NULL, need, d_run_first, d_first_sorted, d_run_index, d_index_sorted, (int)npat, 0, 32, st));
	if (need > tmp_bytes) {
		PAT_CHECK(cudaFree(d_tmp));
		d_tmp = NULL;
		PAT_CHECK(cudaMalloc(&d_tmp, need));
		tmp_bytes = need;
	}
	need = tmp_bytes;
	PAT_CHECK(cub::DeviceRadixSort::SortPairs(d_tmp, need, d_run_first, d_first_sorted, d_run_index, d_index_sorted, (int)npat, 0, 32, st));
	h_hash = (uint32_t *)malloc((size_t)npat * 4);
	h_index = (uint32_t *)malloc((size_t)npat * 4);
	h_hash_q = (uint32_t *)malloc((size_t)npat * 4);
	h_order = (uint32_t *)malloc((size_t)npat * 4);
	if (!h_hash || !h_index || !h_hash_q || !h_order) {
		rc = PHB_ENOMEM;
		goto done;
	}
	PAT_CHECK(cudaMemcpyAsync(h_hash, d_run_hash, (size_t)npat * 4, cudaMemcpyDeviceToHost, st));
	PAT_CHECK(cudaMemcpyAsync(h_index, d_index_sorted, (size_t)npat * 4, cudaMemcpyDeviceToHost, st));
	PAT_CHECK(cudaStreamSynchronize(st));
	for (uint32_t q = 0; q < npat; q++) h_hash_q[q] = h_hash[h_index[q]];  // hash of the q-th new key
	if ((rc = replay_reference_table(npat, h_hash_q, hashtable_size > 0 ? (unsigned)hashtable_size : 100u, h_order)) != PHB_OK) {
		snprintf(pat_err, sizeof(pat_err), "phb_compress_patterns: hash table replay failed");
		goto done;
	}
	for (uint32_t k = 0; k < npat; k++) h_order[k] = h_index[h_order[k]];  // reference position k -> run id
	PAT_CHECK(cudaMemcpyAsync(d_order, h_order, (size_t)npat * 4, cudaMemcpyHostToDevice, st));
	PAT_CHECK(cudaMalloc((void **)&d_pat, (size_t)ntaxa * npat));
	PAT_CHECK(cudaMalloc((void **)&d_w, (size_t)npat * sizeof(double)));
	k_pat_gather<<<(npat + 255) / 256, 256, 0, st>>>(ntaxa, n, npat, d_aln, d_order, d_run_first, d_run_start, d_pat, d_w, d_pos);
	*patterns = (uint8_t *)malloc((size_t)ntaxa * npat);
	*weights = (double *)malloc((size_t)npat * sizeof(double));
	if (!*patterns || !*weights) {
		rc = PHB_ENOMEM;
		goto done;
	}
	PAT_CHECK(cudaMemcpyAsync(*patterns, d_pat, (size_t)ntaxa * npat, cudaMemcpyDeviceToHost, st));
	PAT_CHECK(cudaMemcpyAsync(*weights, d_w, (size_t)npat * sizeof(double), cudaMemcpyDeviceToHost, st));
	if (site_to_pattern) {
		PAT_CHECK(cudaMalloc((void **)&d_map, n * sizeof(int)));
		k_pat_site_map<<<blocks, 256, 0, st>>>(n, d_runid, d_site2, d_pos, d_map);
		*site_to_pattern = (int *)malloc(n * sizeof(int));
		if (!*site_to_pattern) {
			rc = PHB_ENOMEM;
			goto done;
		}
		PAT_CHECK(cudaMemcpyAsync(*site_to_pattern, d_map, n * sizeof(int), cudaMemcpyDeviceToHost, st));
	}
	PAT_CHECK(cudaStreamSynchronize(st));
	PAT_CHECK(cudaGetLastError());
	*npatterns = npat;
done:
	if (rc != PHB_OK) {
		free(*patterns), free(*weights);
		*patterns = NULL, *weights = NULL;
		if (site_to_pattern) {
			free(*site_to_pattern);
			*site_to_pattern = NULL;
		}
	}
	free(h_hash), free(h_index), free(h_hash_q), free(h_order);
	{
		void *bufs[] = {d_aln, d_pat, d_key, d_key2, d_site, d_site2, d_head, d_runid, d_run_first, d_run_hash, d_run_start, d_run_index,
		                d_first_sorted, d_index_sorted, d_order, d_pos, d_w, d_map, d_coll, d_tmp};
		for (size_t i = 0; i < sizeof(bufs) / sizeof(bufs[0]); i++)
			if (bufs[i]) cudaFree(bufs[i]);
	}
	if (st) cudaStreamDestroy(st);
	return rc;
}

extern "C" void phb_free(void *p) { free(p); }
