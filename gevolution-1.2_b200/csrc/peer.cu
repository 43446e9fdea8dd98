// peer.cu -- rank-to-rank communication over peer memory (one node, NVLink 5 / NVSwitch)
//
// Replaces, for ranks that can map each other's memory (cudaIpc), the NCCL point-to-point messages of
//   Field::updateHalo                      (main.cpp:518,568,598)
//   scalar / vector / symtensor *_comm     (gevolution.hpp:1024,1149,1300; main.cpp:411,435,450)
//   the particle hand-over of moveParticles (LATfield2; main.cpp:798), see geodesic.cu
// and the 4-byte all-reduce that served as rank barrier of the FFT exchange (fft.cu).  The messages of these calls
// are a few planes of 2 MB each: their cost over NCCL is launch latency and host round trips (0.6 + 0.74 ms of a
// 9.6 ms cycle on 8 GPUs), not bandwidth.  Here a sender kernel stores straight into the receiver's
// buffer, a flag barrier (one store per peer, one spin per peer, system-scope release / acquire) orders it, and the
// receiver's kernel picks the data up -- three short launches on the library's stream, no host synchronisation.
//
// Every rank owns one communication buffer, mapped by all ranks.  It holds two slots that alternate from one collective
// operation to the next, so that a sender may fill slot s of operation k + 2 only after the barrier of operation k + 1,
// which the receiver enters after it has consumed slot s of operation k (same argument as the two FFT exchange buffers).
//   slot = [ planes region: pc_plane_doubles ][ migration region: 2 directions x (7 x pc_mig_cap doubles) ][ counts ]
#include "gevb_internal.cuh"

namespace {

struct PeerFlags { unsigned long long * p[GEVB_MAX_RANKS]; };

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long * p)
{
	unsigned long long v;
	asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}

// thread r signals rank r (its slot `rank` of r's flag page becomes `epoch`) and waits for r's signal in its own page.
// Everything this rank enqueued before the barrier has completed (stream order); the system-scope release makes it
// visible to the peer before the flag.  A peer that never arrives (it failed) must not hang the GPU: after about two
// seconds the barrier gives up and raises the error flag that the host reads at its next synchronisation.
__global__ void k_peer_barrier(PeerFlags F, int rank, int nranks, unsigned long long epoch, int * err)
{
	const int r = threadIdx.x;
	if (r >= nranks || r == rank) return;
	__threadfence_system();
	atomicMax_system(F.p[r] + rank, epoch);                 // never moves a flag backwards, whatever order two barriers of different streams run in
	const unsigned long long * mine = F.p[rank] + r;
	const long long t0 = clock64();
	while (ld_acquire_sys(mine) < epoch)
		if (clock64() - t0 > 4000000000ll) { *err = 1; break; }
}

// dst[k * n + i] = src[k * stride + i]: boundary planes of every component into a (peer) buffer
__global__ void k_planes_out(double * __restrict__ dst, const double * __restrict__ src, size_t n, int ncomp, size_t stride)
{
	for (int k = 0; k < ncomp; k++)
		for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
			dst[k * n + i] = src[k * stride + i];
}
// dst[k * stride + i] (+)= src[k * n + i]
template <bool ADD>
__global__ void k_planes_in(double * __restrict__ dst, const double * __restrict__ src, size_t n, int ncomp, size_t stride)
{
	for (int k = 0; k < ncomp; k++)
		for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
		{
			const double v = __ldcg(src + k * n + i);       // written by the peer over NVLink: read at L2, never from a stale L1 line
			if (ADD) dst[k * stride + i] += v; else dst[k * stride + i] = v;
		}
}

} // namespace

bool gevb_peer_on(const gevb_ctx * c) { return c->nranks > 1 && c->peer_state == 1 && gevb_tune(TUNE_PEER_COMM) != 0; }

// cudaIpc handle of `mine` to every rank through the communicator; all[r] is rank r's allocation as mapped here.
// Returns 0 with all[] complete, or leaves entries NULL where a mapping failed (the caller agrees on the outcome).
int gevb_peer_share(gevb_ctx * c, void * mine, void ** all, cudaStream_t stream)
{
	cudaIpcMemHandle_t hm, ha[GEVB_MAX_RANKS];
	memset(&hm, 0, sizeof(hm));
	int ok = mine != NULL && cudaIpcGetMemHandle(&hm, mine) == cudaSuccess;
	cudaGetLastError();
	char * stage = (char *) (c->d_red + 5000);                                   // (nranks + 1) * 64 bytes of the reduction buffer
	const size_t hb = sizeof(hm);
	CUDA_TRY(cudaMemcpyAsync(stage + (size_t) c->nranks * hb, &hm, hb, cudaMemcpyHostToDevice, stream));
	NCCL_TRY(ncclGroupStart());
	for (int r = 0; r < c->nranks; r++)
	{
		NCCL_TRY(ncclSend(stage + (size_t) c->nranks * hb, hb, ncclChar, r, c->comm, stream));
		NCCL_TRY(ncclRecv(stage + (size_t) r * hb, hb, ncclChar, r, c->comm, stream));
	}
	NCCL_TRY(ncclGroupEnd());
	CUDA_TRY(cudaMemcpyAsync(ha, stage, (size_t) c->nranks * hb, cudaMemcpyDeviceToHost, stream));
	CUDA_TRY(cudaStreamSynchronize(stream));
	for (int r = 0; r < c->nranks; r++)
	{
		all[r] = NULL;
		if (r == c->rank) { all[r] = mine; continue; }
		if (!ok) continue;
		if (cudaIpcOpenMemHandle(&all[r], ha[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) all[r] = NULL;
	}
	cudaGetLastError();
	return 0;
}

static size_t slot_doubles(const gevb_ctx * c) { return c->pc_plane_doubles + 2 * 7 * c->pc_mig_cap + 16; }

double * gevb_peer_slot(gevb_ctx * c, int rank, int slot) { return (double *) c->pc[rank] + (size_t) slot * slot_doubles(c); }

int gevb_peer_setup(gevb_ctx * c)
{
	if (c->nranks == 1 || c->peer_state != 0) return 0;
	const size_t pl = c->plane();
	// planes region: up to 8 components x 2 planes (halo of a 6-component field and the folds fit);
	// migration: as many particles per direction as half the slab has cells, at least 2^16 (GEVB_MIGRATION_CAP overrides)
	c->pc_plane_doubles = 16 * pl;
	c->pc_mig_cap = (size_t) c->nzl * pl / 2;
	if (c->pc_mig_cap < (1u << 16)) c->pc_mig_cap = 1u << 16;
	if (const char * e = getenv("GEVB_MIGRATION_CAP")) { const long long v = atoll(e); if (v > 0) c->pc_mig_cap = (size_t) v; }
	c->pc_bytes = 2 * slot_doubles(c) * sizeof(double);
	void * buf = NULL;
	unsigned long long * flags = NULL;
	int ok = cudaMalloc(&buf, c->pc_bytes) == cudaSuccess && cudaMalloc(&flags, 4096) == cudaSuccess;
	cudaGetLastError();
	if (ok) { CUDA_TRY(cudaMemsetAsync(flags, 0, 4096, c->stream)); CUDA_TRY(cudaMemsetAsync(buf, 0, c->pc_bytes, c->stream)); }
	CUDA_TRY(cudaMalloc(&c->d_peer_err, 64));
	CUDA_TRY(cudaMemsetAsync(c->d_peer_err, 0, 64, c->stream));
	GEVB_TRY(gevb_peer_share(c, ok ? buf : NULL, c->pc, c->stream));
	GEVB_TRY(gevb_peer_share(c, ok ? (void *) flags : NULL, (void **) c->pf, c->stream));
	c->pc[c->rank] = buf; c->pf[c->rank] = flags;           // owned entries (possibly NULL): release frees them
	for (int r = 0; r < c->nranks; r++) if (c->pc[r] == NULL || c->pf[r] == NULL) ok = 0;
	// agree: one rank without mappings sends everyone to the NCCL messages
	int * flag = (int *) (c->d_red + 5200);
	CUDA_TRY(cudaMemcpyAsync(flag, &ok, sizeof(int), cudaMemcpyHostToDevice, c->stream));
	NCCL_TRY(ncclAllReduce(flag, flag, 1, ncclInt, ncclMin, c->comm, c->stream));
	CUDA_TRY(cudaMemcpyAsync(&ok, flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	if (!ok)
	{
		c->peer_state = 1;                                  // so that release unmaps what was mapped and frees what was allocated
		gevb_peer_release(c);
		c->peer_state = -1;
		fprintf(stderr, "gevb: rank %d: peer memory mapping unavailable, halo / fold / migration go through NCCL\n", c->rank);
		return 0;
	}
	c->peer_state = 1;
	return 0;
}

void gevb_peer_release(gevb_ctx * c)
{
	if (c->peer_state != 1) return;
	for (int r = 0; r < GEVB_MAX_RANKS; r++)
	{
		if (c->pc[r]) { if (r == c->rank) cudaFree(c->pc[r]); else cudaIpcCloseMemHandle(c->pc[r]); c->pc[r] = NULL; }
		if (c->pf[r]) { if (r == c->rank) cudaFree(c->pf[r]); else cudaIpcCloseMemHandle(c->pf[r]); c->pf[r] = NULL; }
	}
	if (c->d_peer_err) { cudaFree(c->d_peer_err); c->d_peer_err = NULL; }
	c->peer_state = 0;
}

int gevb_peer_barrier(gevb_ctx * c, cudaStream_t stream)
{
	PeerFlags F;
	for (int r = 0; r < GEVB_MAX_RANKS; r++) F.p[r] = c->pf[r];
	k_peer_barrier<<<1, 32, 0, stream>>>(F, c->rank, c->nranks, ++c->pf_epoch, c->d_peer_err);
	KERNEL_CHECK(c);
	return 0;
}

int gevb_peer_check(gevb_ctx * c)
{
	if (c->nranks == 1 || c->peer_state != 1) return 0;
	int err = 0;
	CUDA_TRY(cudaMemcpy(&err, c->d_peer_err, sizeof(int), cudaMemcpyDeviceToHost));
	GEVB_CHECK_ARG(err == 0, "a rank barrier over peer memory timed out (a peer rank failed or left the collective sequence)");
	return 0;
}

// Field::updateHalo: my first bulk plane becomes the lower neighbour's upper ghost plane, my last bulk plane the
// upper neighbour's lower ghost plane
int gevb_peer_halo(gevb_field * f)
{
	gevb_ctx * c = f->ctx;
	const size_t pl = c->plane(), n = pl;
	const int up = (c->rank + 1) % c->nranks, dn = (c->rank + c->nranks - 1) % c->nranks;
	const int slot = (int) (c->pc_seq++ & 1);
	const int grid = gevb_grid(c, n, 256, 4);
	// receiver's planes region: [0, ncomp * pl) what arrives from above, [ncomp * pl, 2 ncomp * pl) what arrives from below
	k_planes_out<<<grid, 256, 0, c->stream>>>(gevb_peer_slot(c, dn, slot), f->data + pl, n, f->ncomp, f->comp_stride);
	KERNEL_CHECK(c);
	k_planes_out<<<grid, 256, 0, c->stream>>>(gevb_peer_slot(c, up, slot) + (size_t) f->ncomp * pl, f->data + (size_t) c->nzl * pl, n, f->ncomp, f->comp_stride);
	KERNEL_CHECK(c);
	GEVB_TRY(gevb_peer_barrier(c, c->stream));
	const double * mine = gevb_peer_slot(c, c->rank, slot);
	k_planes_in<false><<<grid, 256, 0, c->stream>>>(f->data + (size_t) (c->nzl + 1) * pl, mine, n, f->ncomp, f->comp_stride);
	KERNEL_CHECK(c);
	k_planes_in<false><<<grid, 256, 0, c->stream>>>(f->data, mine + (size_t) f->ncomp * pl, n, f->ncomp, f->comp_stride);
	KERNEL_CHECK(c);
	return 0;
}

// *_comm: my upper ghost plane is added into the upper neighbour's first bulk plane
int gevb_peer_fold(gevb_field * f)
{
	gevb_ctx * c = f->ctx;
	const size_t pl = c->plane(), n = pl;
	const int up = (c->rank + 1) % c->nranks;
	const int slot = (int) (c->pc_seq++ & 1);
	const int grid = gevb_grid(c, n, 256, 4);
	k_planes_out<<<grid, 256, 0, c->stream>>>(gevb_peer_slot(c, up, slot), f->data + (size_t) (c->nzl + 1) * pl, n, f->ncomp, f->comp_stride);
	KERNEL_CHECK(c);
	GEVB_TRY(gevb_peer_barrier(c, c->stream));
	k_planes_in<true><<<grid, 256, 0, c->stream>>>(f->data + pl, gevb_peer_slot(c, c->rank, slot), n, f->ncomp, f->comp_stride);
	KERNEL_CHECK(c);
	return 0;
}
