// fft.cu -- PlanFFT<Cplx>::execute (main.cpp:238-246,477,488,544,563,575,593)
//
// Per-component 3-D r2c (forward, e^{-ikx}) / c2r (backward), unnormalised both
// ways (manual.pdf section 4).  cuFFT does the local transforms; with more than
// one rank the global transpose is one NCCL all-to-all per execute:
//
//   nranks == 1 : one batched 3-D D2Z / Z2D plan over all components.
//   nranks  > 1 : forward  = 2-D D2Z on every local z-plane, written by cuFFT's
//                            strided output directly in send order [ky][zl][kx]
//                            (the ky-slab of every destination rank is contiguous)
//                            -> all-to-all -> tile transpose to [ky][kx][kz]
//                            -> 1-D Z2Z along kz
//                 backward = 1-D Z2Z^-1 along kz (out of place) -> tile transpose
//                            into send order -> all-to-all, which lands as
//                            [ky][zl][kx] -> 2-D Z2D per plane reading that layout.
//                 One pass over the data less per direction than pack / exchange /
//                 unpack: the exchange buffers are cuFFT's own output / input.
//
// Exchange over peer memory (default when the ranks can map each other's memory, cudaIpc over NVLink / NVSwitch):
// the transpose kernel itself performs the exchange -- it reads local tiles and stores the transposed tiles
// straight into the destination rank's buffer, already in the layout the next local transform reads, so the
// exchange and the layout change are one pass and the NVLink traffic of all peers overlaps tile by tile:
//
//                 forward  = 2-D D2Z (local)  ->  k_push_fwd: [ky][zl][kx] tiles -> peer [kyl][kx][kz]  -> barrier
//                            -> 1-D Z2Z out of place from the exchange buffer into the Fourier field
//                 backward = 1-D Z2Z^-1 (local, out of place) -> k_push_bwd: [kx][z] tiles -> peer [ky][zl][kx]
//                            -> barrier -> 2-D Z2D from the exchange buffer
//
// Two exchange buffers alternate, so one barrier per transform (a 4-byte NCCL all-reduce on the same stream) is
// enough: nobody can start overwriting a buffer before every rank has passed the barrier of the transform in
// between, which each rank enters only after it has consumed the buffer.  tuning knob fft_exchange = 0 selects
// the NCCL path above (also the fallback where peer mapping is unavailable).
//
// The Fourier input of a backward transform is preserved (the reference keeps
// BiFT as persistent state across steps, main.cpp:586-593): cuFFT's multi-dim
// Z2D may overwrite its input, so on one rank the z-pass runs out of place into
// scratch (1-D Z2Z with stride N*nh) and the 2-D Z2D of the planes consumes the
// scratch -- three passes as in the 3-D plan, no copy of the field.
#include "gevb_internal.cuh"

namespace {

// recv buffer [src rank][c][kyl][zl of src][kx]  ->  slab layout [c][kyl][kx][kz = src*nzl + zl]
// tile transpose (kz <-> kx) through shared memory so that reads and writes are both coalesced
__global__ void k_unpack_fwd(const double2 * __restrict__ recv, double2 * __restrict__ out, int N, int nh, int nzl, int nkyl, int ncomp, int nranks)
{
	__shared__ double2 tile[32][33];
	// grid: x = kx tiles, y = kz tiles, z = c * nkyl + kyl
	const int c = blockIdx.z / nkyl, kyl = blockIdx.z % nkyl;
	const int kx0 = blockIdx.x * 32, kz0 = blockIdx.y * 32;
	for (int j = threadIdx.y; j < 32; j += blockDim.y)
	{
		int kz = kz0 + j, kx = kx0 + threadIdx.x;
		if (kz < N && kx < nh)
		{
			int src = kz / nzl, zl = kz % nzl;
			tile[j][threadIdx.x] = recv[((((size_t) src * ncomp + c) * nkyl + kyl) * nzl + zl) * nh + kx];
		}
	}
	__syncthreads();
	for (int j = threadIdx.y; j < 32; j += blockDim.y)
	{
		int kx = kx0 + j, kz = kz0 + threadIdx.x;
		if (kz < N && kx < nh) out[(((size_t) c * nkyl + kyl) * nh + kx) * N + kz] = tile[threadIdx.x][j];
	}
}

// slab layout [c][kyl][kx][z]  ->  send buffer [dst rank][c][kyl][zl in dst slab][kx]   (transpose back)
__global__ void k_pack_bwd(const double2 * __restrict__ in, double2 * __restrict__ send, int N, int nh, int nzl, int nkyl, int ncomp, int nranks)
{
	__shared__ double2 tile[32][33];
	const int c = blockIdx.z / nkyl, kyl = blockIdx.z % nkyl;
	const int kx0 = blockIdx.x * 32, z0 = blockIdx.y * 32;
	for (int j = threadIdx.y; j < 32; j += blockDim.y)
	{
		int kx = kx0 + j, z = z0 + threadIdx.x;
		if (z < N && kx < nh) tile[j][threadIdx.x] = in[(((size_t) c * nkyl + kyl) * nh + kx) * N + z];
	}
	__syncthreads();
	for (int j = threadIdx.y; j < 32; j += blockDim.y)
	{
		int z = z0 + j, kx = kx0 + threadIdx.x;
		if (z < N && kx < nh)
		{
			int dst = z / nzl, zl = z % nzl;
			send[((((size_t) dst * ncomp + c) * nkyl + kyl) * nzl + zl) * nh + kx] = tile[threadIdx.x][j];
		}
	}
}

struct PeerPtrs { double2 * p[GEVB_MAX_RANKS]; };

// forward exchange: local A [c][ky][zl][kx]  ->  rank d = ky / nkyl : X_d [c][kyl][kx][kz = rank * nzl + zl]
// Blocks walk the 32 x 32 tiles with stride gridDim.x: the launch decides how much of the machine the exchange takes
// (all of it when it runs alone, a few hundred resident blocks when it shares the SMs with the next local transform).
// (zb, zn): only the local planes [zb, zb + zn) -- one chunk of the plane pipeline (zn is a multiple of 32 or the whole slab)
__global__ void __launch_bounds__(256) k_push_fwd(const double2 * __restrict__ A, PeerPtrs X, int N, int nh, int nzl, int nkyl, int rank, int c0, int ncomp, int zb, int zn)
{
	__shared__ double2 tile[32][33];
	const int tkx = (nh + 31) / 32, tz = (zn + 31) / 32;
	const long ntiles = (long) tkx * tz * ncomp * N;
	for (long t = blockIdx.x; t < ntiles; t += gridDim.x)
	{
		const int kx0 = (int) (t % tkx) * 32; long r = t / tkx;
		const int zl0 = zb + (int) (r % tz) * 32; r /= tz;
		const int c = c0 + (int) (r / N), ky = (int) (r % N);
		const int d = ky / nkyl, kyl = ky % nkyl;
		const double2 * src = A + ((size_t) c * N + ky) * nzl * nh;
		for (int j = threadIdx.y; j < 32; j += blockDim.y)
		{
			const int zl = zl0 + j, kx = kx0 + threadIdx.x;
			if (zl < zb + zn && kx < nh) tile[j][threadIdx.x] = __ldcs(src + (size_t) zl * nh + kx);
		}
		__syncthreads();
		double2 * dst = X.p[d] + ((size_t) c * nkyl + kyl) * nh * N + (size_t) rank * nzl;
		for (int j = threadIdx.y; j < 32; j += blockDim.y)
		{
			const int kx = kx0 + j, zl = zl0 + threadIdx.x;
			if (zl < zb + zn && kx < nh) dst[(size_t) kx * N + zl] = tile[threadIdx.x][j];
		}
		__syncthreads();
	}
}

// backward exchange: local A [c][kyl][kx][z]  ->  rank d = z / nzl : X_d [c][ky = rank * nkyl + kyl][zl][kx]
// (yb, yn): only the local rows kyl in [yb, yb + yn) -- one chunk of the row pipeline
__global__ void __launch_bounds__(256) k_push_bwd(const double2 * __restrict__ A, PeerPtrs X, int N, int nh, int nzl, int nkyl, int rank, int c0, int ncomp, int yb, int yn)
{
	__shared__ double2 tile[32][33];
	const int tkx = (nh + 31) / 32, tz = (N + 31) / 32;
	const long ntiles = (long) tkx * tz * ncomp * yn;
	for (long t = blockIdx.x; t < ntiles; t += gridDim.x)
	{
		const int kx0 = (int) (t % tkx) * 32; long r = t / tkx;
		const int z0 = (int) (r % tz) * 32; r /= tz;
		const int c = c0 + (int) (r / yn), kyl = yb + (int) (r % yn);
		const double2 * src = A + ((size_t) c * nkyl + kyl) * nh * N;
		for (int j = threadIdx.y; j < 32; j += blockDim.y)
		{
			const int kx = kx0 + j, z = z0 + threadIdx.x;
			if (z < N && kx < nh) tile[j][threadIdx.x] = __ldcs(src + (size_t) kx * N + z);
		}
		__syncthreads();
		const size_t row = ((size_t) c * N + (size_t) rank * nkyl + kyl) * nzl;
		for (int j = threadIdx.y; j < 32; j += blockDim.y)
		{
			const int z = z0 + j, kx = kx0 + threadIdx.x;
			if (z < N && kx < nh)
			{
				const int d = z / nzl, zl = z % nzl;
				X.p[d][(row + zl) * nh + kx] = tile[threadIdx.x][j];
			}
		}
		__syncthreads();
	}
}

// stream-ordered barrier across the ranks: every rank's earlier work on its stream (the pushes into peer memory) has
// completed before any rank's later work starts
int rank_barrier(gevb_ctx * c, cudaStream_t stream)
{
	if (gevb_peer_on(c)) return gevb_peer_barrier(c, stream);       // flags in peer memory: one 32-thread kernel instead of a collective
	NCCL_TRY(ncclAllReduce(c->d_barrier, c->d_barrier, 1, ncclInt, ncclSum, c->comm, stream));
	return 0;
}

// (re)allocate the two exchange buffers and map every peer's pair (collective: all ranks create the same plans in the
// same order).  Any failure to map leaves xchg_state = -1 on every rank and the NCCL exchange in use.
int xchg_ensure(gevb_ctx * c, size_t bytes)
{
	if (c->xchg_state < 0 || (c->xchg_state == 1 && c->xchg_bytes >= bytes)) return 0;
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	gevb_xchg_release(c);
	if (c->d_barrier == NULL)
	{
		CUDA_TRY(cudaMalloc(&c->d_barrier, 64)); CUDA_TRY(cudaMemset(c->d_barrier, 0, 64));
		int prio_lo = 0, prio_hi = 0;
		CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
		CUDA_TRY(cudaStreamCreateWithPriority(&c->xstream, cudaStreamNonBlocking, prio_hi));     // its blocks are placed before the local transform's
		for (int k = 0; k <= GEVB_NXEV; k++) CUDA_TRY(cudaEventCreateWithFlags(&c->xev[k], cudaEventDisableTiming));
	}
	int ok = 1;
	cudaIpcMemHandle_t mine[2], all[GEVB_MAX_RANKS][2];
	memset(mine, 0, sizeof(mine));
	for (int b = 0; b < 2 && ok; b++)
	{
		if (cudaMalloc(&c->xchg[b][c->rank], bytes) != cudaSuccess) { c->xchg[b][c->rank] = NULL; ok = 0; break; }
		if (cudaIpcGetMemHandle(&mine[b], c->xchg[b][c->rank]) != cudaSuccess) ok = 0;
	}
	cudaGetLastError();
	// all handles to all ranks through the communicator (device staging in the reduction buffer)
	char * stage = (char *) (c->d_red + 5000);                                   // (nranks + 1) * 128 bytes
	const size_t hb = sizeof(mine);
	CUDA_TRY(cudaMemcpyAsync(stage + (size_t) c->nranks * hb, mine, hb, cudaMemcpyHostToDevice, c->stream));
	NCCL_TRY(ncclGroupStart());
	for (int r = 0; r < c->nranks; r++)
	{
		NCCL_TRY(ncclSend(stage + (size_t) c->nranks * hb, hb, ncclChar, r, c->comm, c->stream));
		NCCL_TRY(ncclRecv(stage + (size_t) r * hb, hb, ncclChar, r, c->comm, c->stream));
	}
	NCCL_TRY(ncclGroupEnd());
	CUDA_TRY(cudaMemcpyAsync(all, stage, (size_t) c->nranks * hb, cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	for (int r = 0; r < c->nranks && ok; r++)
	{
		if (r == c->rank) continue;
		for (int b = 0; b < 2; b++)
			if (cudaIpcOpenMemHandle(&c->xchg[b][r], all[r][b], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { c->xchg[b][r] = NULL; ok = 0; }
	}
	cudaGetLastError();
	// agree: one rank without mappings sends everyone to the NCCL exchange
	int * flag = c->d_barrier + 1;
	CUDA_TRY(cudaMemcpyAsync(flag, &ok, sizeof(int), cudaMemcpyHostToDevice, c->stream));
	NCCL_TRY(ncclAllReduce(flag, flag, 1, ncclInt, ncclMin, c->comm, c->stream));
	CUDA_TRY(cudaMemcpyAsync(&ok, flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	if (!ok)
	{
		gevb_xchg_release(c);
		c->xchg_state = -1;
		fprintf(stderr, "gevb: rank %d: peer memory mapping unavailable, slab FFT exchanges through NCCL\n", c->rank);
		return 0;
	}
	c->xchg_state = 1; c->xchg_bytes = bytes;
	return 0;
}

// one chunk of `chunk` complex numbers per (peer rank, component).  Send side: [peer][c] when send_peer_major, else
// [c][peer]; same for the receive side.
int alltoall(gevb_ctx * c, const double2 * send, double2 * recv, size_t chunk, int ncomp, bool send_peer_major, bool recv_peer_major)
{
	NCCL_TRY(ncclGroupStart());
	for (int r = 0; r < c->nranks; r++)
		for (int k = 0; k < ncomp; k++)
		{
			const size_t so = send_peer_major ? (size_t) r * ncomp + k : (size_t) k * c->nranks + r;
			const size_t ro = recv_peer_major ? (size_t) r * ncomp + k : (size_t) k * c->nranks + r;
			NCCL_TRY(ncclSend(send + so * chunk, chunk * 2, ncclDouble, r, c->comm, c->stream));
			NCCL_TRY(ncclRecv(recv + ro * chunk, chunk * 2, ncclDouble, r, c->comm, c->stream));
		}
	NCCL_TRY(ncclGroupEnd());
	return 0;
}

} // namespace

void gevb_xchg_release(gevb_ctx * c)
{
	for (int b = 0; b < 2; b++)
		for (int r = 0; r < GEVB_MAX_RANKS; r++)
		{
			if (c->xchg[b][r] == NULL) continue;
			if (r == c->rank) cudaFree(c->xchg[b][r]); else cudaIpcCloseMemHandle(c->xchg[b][r]);
			c->xchg[b][r] = NULL;
		}
	c->xchg_bytes = 0;
	if (c->xchg_state == 1) c->xchg_state = 0;
}

// tuning knob fft_l2_planes = L > 0 (single rank): the 2-D transforms run L planes at a time, so that cuFFT's x-pass and
// y-pass of a chunk follow each other while the chunk (L N^2 reals in, L N nh complex out) is still in the 126 MB L2 and
// the intermediate never travels to HBM; 0: all planes in one batched call (two HBM passes).  Returns the chunk in use.
static int chunked_planes(gevb_plan * p)
{
	const int L = gevb_tune(TUNE_FFT_L2_PLANES);
	const int N = p->ctx->N, nh = p->ctx->nh;
	if (L <= 0 || L >= N || N % L != 0) return 0;
	if (p->chunk_planes != L)
	{
		if (p->chunk_planes) { cufftDestroy(p->f2d_c); cufftDestroy(p->b2d_c); p->chunk_planes = 0; }
		int n2[2] = {N, N}, r2[2] = {N, N}, k2[2] = {N, nh};
		if (cufftPlanMany(&p->f2d_c, 2, n2, r2, 1, N * N, k2, 1, N * nh, CUFFT_D2Z, L) != CUFFT_SUCCESS) return 0;
		if (cufftPlanMany(&p->b2d_c, 2, n2, k2, 1, N * nh, r2, 1, N * N, CUFFT_Z2D, L) != CUFFT_SUCCESS) { cufftDestroy(p->f2d_c); return 0; }
		cufftSetStream(p->f2d_c, p->ctx->stream); cufftSetStream(p->b2d_c, p->ctx->stream);
		p->chunk_planes = L;
	}
	return L;
}

extern "C" int gevb_plan_create(gevb_plan ** out, gevb_field * rf, gevb_field * cf)
{
	GEVB_CHECK_ARG(out != NULL && rf != NULL && cf != NULL, "gevb_plan_create: NULL argument");
	GEVB_CHECK_ARG(rf->kind == GEVB_REAL && cf->kind == GEVB_CPLX, "gevb_plan_create: needs (real field, Fourier field)");
	GEVB_CHECK_ARG(rf->ncomp == cf->ncomp, "gevb_plan_create: component counts differ (%d vs %d)", rf->ncomp, cf->ncomp);
	GEVB_CHECK_ARG(rf->ctx == cf->ctx, "gevb_plan_create: fields belong to different contexts");
	gevb_ctx * c = rf->ctx;
	CUDA_TRY(cudaSetDevice(c->device));
	gevb_plan * p = new gevb_plan();
	memset(p, 0, sizeof(*p));
	p->ctx = c; p->real_field = rf; p->cplx_field = cf; p->multi = c->nranks > 1; p->preserve = true;
	const int N = c->N, nh = c->nh;
	if (!p->multi)
	{
		int n[3] = {N, N, N};
		int rembed[3] = {N, N, N}, kembed[3] = {N, N, nh};
		CUFFT_TRY(cufftPlanMany(&p->fwd, 3, n, rembed, 1, (int) rf->comp_stride, kembed, 1, (int) cf->comp_stride, CUFFT_D2Z, rf->ncomp));
		CUFFT_TRY(cufftPlanMany(&p->bwd, 3, n, kembed, 1, (int) cf->comp_stride, rembed, 1, (int) rf->comp_stride, CUFFT_Z2D, rf->ncomp));
		CUFFT_TRY(cufftSetStream(p->fwd, c->stream));
		CUFFT_TRY(cufftSetStream(p->bwd, c->stream));
		// input-preserving backward transform in the same three passes, without a copy of the Fourier field: the z-pass runs
		// out of place into scratch (lines along kz have stride N*nh, consecutive lines are consecutive complex numbers),
		// the 2-D Z2D of every z-plane then may clobber the scratch
		int n1[1] = {N}, n2[2] = {N, N};
		int e1[1] = {N}, r2[2] = {N, N}, k2[2] = {N, nh};
		CUFFT_TRY(cufftPlanMany(&p->bz1d, 1, n1, e1, N * nh, 1, e1, N * nh, 1, CUFFT_Z2Z, N * nh));
		CUFFT_TRY(cufftPlanMany(&p->b2d, 2, n2, k2, 1, N * nh, r2, 1, N * N, CUFFT_Z2D, N));
		CUFFT_TRY(cufftPlanMany(&p->f2d, 2, n2, r2, 1, N * N, k2, 1, N * nh, CUFFT_D2Z, N));
		CUFFT_TRY(cufftSetStream(p->f2d, c->stream));
		CUFFT_TRY(cufftSetStream(p->bz1d, c->stream));
		CUFFT_TRY(cufftSetStream(p->b2d, c->stream));
	}
	else
	{
		int n2[2] = {N, N};
		// real side: plane after plane; Fourier side: [ky][zl][kx], i.e. row pitch nzl*nh and plane distance nh
		int rembed[2] = {N, N}, kembed[2] = {N, c->nzl * nh};
		CUFFT_TRY(cufftPlanMany(&p->fwd2d, 2, n2, rembed, 1, N * N, kembed, 1, nh, CUFFT_D2Z, c->nzl));
		CUFFT_TRY(cufftPlanMany(&p->bwd2d, 2, n2, kembed, 1, nh, rembed, 1, N * N, CUFFT_Z2D, c->nzl));
		int n1[1] = {N};
		CUFFT_TRY(cufftPlanMany(&p->z1d, 1, n1, n1, 1, N, n1, 1, N, CUFFT_Z2Z, rf->ncomp * c->nkyl * nh));
		CUFFT_TRY(cufftSetStream(p->fwd2d, c->stream));
		CUFFT_TRY(cufftSetStream(p->bwd2d, c->stream));
		CUFFT_TRY(cufftSetStream(p->z1d, c->stream));
		CUFFT_TRY(cufftPlanMany(&p->z1d_one, 1, n1, n1, 1, N, n1, 1, N, CUFFT_Z2Z, c->nkyl * nh));
		CUFFT_TRY(cufftSetStream(p->z1d_one, c->stream));
		// pipeline inside a component: the slab's planes (forward) / rows (backward) are transformed in `chunks` pieces and the
		// push of a piece overlaps the local transform of the next, so that a one-component transform hides its exchange too
		// (forward pieces are whole 32-plane tiles of the transposing push; backward pieces are rows of the ky-slab)
		p->chunks = p->chunks_bwd = 1;
		for (int ch = 4; ch > 1; ch >>= 1)
			if (c->nzl % (32 * ch) == 0) { p->chunks = ch; break; }
		for (int ch = 4; ch > 1; ch >>= 1)
			if (c->nkyl % ch == 0 && c->nkyl / ch >= 8) { p->chunks_bwd = ch; break; }
		if (p->chunks > 1)
		{
			CUFFT_TRY(cufftPlanMany(&p->fwd2d_c, 2, n2, rembed, 1, N * N, kembed, 1, nh, CUFFT_D2Z, c->nzl / p->chunks));
			CUFFT_TRY(cufftSetStream(p->fwd2d_c, c->stream));
		}
		if (p->chunks_bwd > 1)
		{
			CUFFT_TRY(cufftPlanMany(&p->z1d_c, 1, n1, n1, 1, N, n1, 1, N, CUFFT_Z2Z, (c->nkyl / p->chunks_bwd) * nh));
			CUFFT_TRY(cufftSetStream(p->z1d_c, c->stream));
		}
		// exchange buffers large enough for this field (collective; grow-only)
		// (sized for six components from the start -- the largest field of the time loop -- so that they are mapped once)
		GEVB_TRY(xchg_ensure(c, (size_t) (rf->ncomp > 6 ? rf->ncomp : 6) * c->nzl * N * nh * sizeof(double2)));
	}
	*out = p;
	return 0;
}

extern "C" int gevb_plan_destroy(gevb_plan * p)
{
	if (p == NULL) return 0;
	cudaSetDevice(p->ctx->device);
	cudaStreamSynchronize(p->ctx->stream);
	if (!p->multi) { cufftDestroy(p->fwd); cufftDestroy(p->bwd); cufftDestroy(p->f2d); cufftDestroy(p->bz1d); cufftDestroy(p->b2d); if (p->chunk_planes) { cufftDestroy(p->f2d_c); cufftDestroy(p->b2d_c); } if (p->yz2d) cufftDestroy(p->yz2d); }
	else { cufftDestroy(p->fwd2d); cufftDestroy(p->bwd2d); cufftDestroy(p->z1d); cufftDestroy(p->z1d_one); if (p->chunks > 1) cufftDestroy(p->fwd2d_c); if (p->chunks_bwd > 1) cufftDestroy(p->z1d_c); }
	delete p;
	return 0;
}

extern "C" int gevb_plan_set_preserve_input(gevb_plan * p, int preserve)
{
	GEVB_CHECK_ARG(p != NULL, "gevb_plan_set_preserve_input: NULL plan");
	p->preserve = preserve != 0;
	return 0;
}

extern "C" int gevb_plan_execute(gevb_plan * p, int direction)
{
	GEVB_CHECK_ARG(p != NULL, "gevb_plan_execute: NULL plan");
	GEVB_CHECK_ARG(direction == GEVB_FFT_FORWARD || direction == GEVB_FFT_BACKWARD, "gevb_plan_execute: bad direction %d", direction);
	gevb_ctx * c = p->ctx;
	gevb_field * rf = p->real_field, * cf = p->cplx_field;
	CUDA_TRY(cudaSetDevice(c->device));
	Timed timed_(c, direction == GEVB_FFT_FORWARD ? CLS_FFT_FWD : CLS_FFT_BWD);
	const int N = c->N, nh = c->nh, nc = rf->ncomp;
	double * rbulk = rf->data + c->plane();
	if (!p->multi)
	{
		// the same three passes as cuFFT's 3-D plan, issued as 2-D per plane + 1-D along z: measured 12 % faster per
		// component (profiles/r1u), and the backward transform can keep its input without a copy
		const bool decomposed = gevb_tune(TUNE_FFT_DECOMPOSED) != 0;
		if (direction == GEVB_FFT_FORWARD)
		{
			if (!decomposed) { CUFFT_TRY(cufftExecD2Z(p->fwd, rbulk, (cufftDoubleComplex *) cf->data)); c->launches++; return 0; }
			if (gevb_xpass_available(p)) return gevb_xpass_forward(p, 0, NULL, NULL, 0., 0., 0., 0., NULL);   // own x-pass + one strided 2-D pass (xpass.cu)
			const int L = chunked_planes(p);
			for (int k = 0; k < nc; k++)
			{
				cufftDoubleComplex * out = (cufftDoubleComplex *) cf->data + k * cf->comp_stride;
				if (L == 0) CUFFT_TRY(cufftExecD2Z(p->f2d, rbulk + k * rf->comp_stride, out));
				else
					for (int z = 0; z < N; z += L)          // x-pass and y-pass of L planes back to back: the intermediate stays in L2
						CUFFT_TRY(cufftExecD2Z(p->f2d_c, rbulk + k * rf->comp_stride + (size_t) z * N * N, out + (size_t) z * N * nh));
				CUFFT_TRY(cufftExecZ2Z(p->bz1d, out, out, CUFFT_FORWARD));
			}
			c->launches += (L ? N / L + 1 : 2) * nc;
		}
		else
		{
			if (!p->preserve)
			{
				// the Fourier field is scratch for the caller
				if (!decomposed) { CUFFT_TRY(cufftExecZ2D(p->bwd, (cufftDoubleComplex *) cf->data, rbulk)); c->launches++; return 0; }
				const int L = chunked_planes(p);
				for (int k = 0; k < nc; k++)
				{
					cufftDoubleComplex * in = (cufftDoubleComplex *) cf->data + k * cf->comp_stride;
					CUFFT_TRY(cufftExecZ2Z(p->bz1d, in, in, CUFFT_INVERSE));
					if (L == 0) CUFFT_TRY(cufftExecZ2D(p->b2d, in, rbulk + k * rf->comp_stride));
					else
						for (int z = 0; z < N; z += L)
							CUFFT_TRY(cufftExecZ2D(p->b2d_c, in + (size_t) z * N * nh, rbulk + k * rf->comp_stride + (size_t) z * N * N));
				}
				c->launches += (L ? N / L + 1 : 2) * nc;
			}
			else
			{
				void * stage;
				GEVB_TRY(gevb_ctx_scratch2(c, cf->bytes, &stage));
				const int L = chunked_planes(p);
				for (int k = 0; k < nc; k++)
				{
					cufftDoubleComplex * in = (cufftDoubleComplex *) cf->data + k * cf->comp_stride, * tmp = (cufftDoubleComplex *) stage + k * cf->comp_stride;
					CUFFT_TRY(cufftExecZ2Z(p->bz1d, in, tmp, CUFFT_INVERSE));
					if (L == 0) CUFFT_TRY(cufftExecZ2D(p->b2d, tmp, rbulk + k * rf->comp_stride));
					else
						for (int z = 0; z < N; z += L)
							CUFFT_TRY(cufftExecZ2D(p->b2d_c, tmp + (size_t) z * N * nh, rbulk + k * rf->comp_stride + (size_t) z * N * N));
				}
				c->launches += (L ? N / L + 1 : 2) * nc;
			}
		}
		return 0;
	}
	// ---- slab-decomposed transform ------------------------------------------------
	const size_t comp_sites = (size_t) c->nzl * N * nh;                 // == nkyl * nh * N: complex sites of one component on this rank
	const size_t wsites = (size_t) nc * comp_sites;
	const size_t chunk = (size_t) c->nkyl * c->nzl * nh;               // what one rank sends to one peer per component
	void * s1, * s2;
	GEVB_TRY(gevb_ctx_scratch(c, wsites * sizeof(double2), &s1));
	GEVB_TRY(gevb_ctx_scratch2(c, wsites * sizeof(double2), &s2));
	double2 * A = (double2 *) s1, * B = (double2 *) s2;
	dim3 tb(32, 8), tg((nh + 31) / 32, (N + 31) / 32, nc * c->nkyl);
	if (c->xchg_state == 1 && gevb_tune(TUNE_FFT_EXCHANGE) != 0)
	{
		// ---- exchange fused into the transpose, over peer memory ----------------------------------
		const int buf = (int) (c->xchg_epoch++ & 1);
		PeerPtrs X;
		for (int r = 0; r < GEVB_MAX_RANKS; r++) X.p[r] = (double2 *) c->xchg[buf][r];
		double2 * Xl = (double2 *) c->xchg[buf][c->rank];
		// component pipeline: the push of component k (xstream) overlaps the local transform of component k+1 (stream);
		// the time the main stream then still waits for the exchange is what CLS_FFT_A2A measures
		// fft_overlap: 0 off, 1 component pipeline, 2 (default) also pieces inside a component for forward transforms, 3 for backward ones
		// too (measured at 8 ranks: pieces gain 0.13 ms per cycle forward and lose 0.17 ms backward, where the z-pass comes first and its
		// pieces leave too little behind them to hide a push -- profiles/round2_v10_multi/ablations_8.jsonl)
		const int ov = gevb_tune(TUNE_FFT_OVERLAP);
		const int chunks = direction == GEVB_FFT_FORWARD ? (ov >= 2 ? p->chunks : 1) : (ov >= 3 ? p->chunks_bwd : 1);
		const bool overlap = gevb_tune(TUNE_FFT_OVERLAP) != 0 && (nc > 1 || chunks > 1) && nc <= 7;
		cudaStream_t xs = overlap ? c->xstream : c->stream;
		// resident blocks of the exchange: two per SM when it shares the machine with a local transform, else eight
		const long pcap = (long) c->num_sms * (overlap ? 2 : 8);
		if (direction == GEVB_FFT_FORWARD)
		{
			const int zn = c->nzl / (overlap ? chunks : 1);
			const long ntiles = (long) ((nh + 31) / 32) * ((zn + 31) / 32) * (overlap ? 1 : nc) * N;
			const int pg = (int) (ntiles < pcap ? ntiles : pcap);
			for (int k = 0; k < nc; k++)
			{
				if (!overlap || chunks == 1)
				{
					CUFFT_TRY(cufftExecD2Z(p->fwd2d, rbulk + k * rf->comp_stride, (cufftDoubleComplex *) (A + (size_t) k * comp_sites)));
					c->launches++;
				}
				for (int ch = 0; overlap && ch < chunks; ch++)
				{
					if (chunks > 1)
					{
						CUFFT_TRY(cufftExecD2Z(p->fwd2d_c, rbulk + k * rf->comp_stride + (size_t) ch * zn * N * N, (cufftDoubleComplex *) (A + (size_t) k * comp_sites + (size_t) ch * zn * nh)));
						c->launches++;
					}
					cudaEvent_t ev = c->xev[(k * chunks + ch) % GEVB_NXEV];
					CUDA_TRY(cudaEventRecord(ev, c->stream));
					CUDA_TRY(cudaStreamWaitEvent(xs, ev, 0));
					k_push_fwd<<<pg, tb, 0, xs>>>(A, X, N, nh, c->nzl, c->nkyl, c->rank, k, 1, ch * zn, zn);
					KERNEL_CHECK(c);
				}
			}
			{
				Timed t_(c, CLS_FFT_A2A);
				if (!overlap) { k_push_fwd<<<pg, tb, 0, xs>>>(A, X, N, nh, c->nzl, c->nkyl, c->rank, 0, nc, 0, c->nzl); KERNEL_CHECK(c); }
				GEVB_TRY(rank_barrier(c, xs));
				if (overlap) { CUDA_TRY(cudaEventRecord(c->xev[GEVB_NXEV], xs)); CUDA_TRY(cudaStreamWaitEvent(c->stream, c->xev[GEVB_NXEV], 0)); }
			}
			CUFFT_TRY(cufftExecZ2Z(p->z1d, (cufftDoubleComplex *) Xl, (cufftDoubleComplex *) cf->data, CUFFT_FORWARD));
			c->launches++;
		}
		else
		{
			const int yn = c->nkyl / (overlap ? chunks : 1);
			const long ntiles = (long) ((nh + 31) / 32) * ((N + 31) / 32) * (overlap ? 1 : nc) * yn;
			const int pg = (int) (ntiles < pcap ? ntiles : pcap);
			if (overlap)
				for (int k = 0; k < nc; k++)
					for (int ch = 0; ch < chunks; ch++)
					{
						const size_t off = (size_t) k * comp_sites + (size_t) ch * yn * nh * N;
						CUFFT_TRY(cufftExecZ2Z(chunks > 1 ? p->z1d_c : p->z1d_one, (cufftDoubleComplex *) cf->data + off, (cufftDoubleComplex *) (A + off), CUFFT_INVERSE));
						c->launches++;
						cudaEvent_t ev = c->xev[(k * chunks + ch) % GEVB_NXEV];
						CUDA_TRY(cudaEventRecord(ev, c->stream));
						CUDA_TRY(cudaStreamWaitEvent(xs, ev, 0));
						k_push_bwd<<<pg, tb, 0, xs>>>(A, X, N, nh, c->nzl, c->nkyl, c->rank, k, 1, ch * yn, yn);
						KERNEL_CHECK(c);
					}
			else
			{
				CUFFT_TRY(cufftExecZ2Z(p->z1d, (cufftDoubleComplex *) cf->data, (cufftDoubleComplex *) A, CUFFT_INVERSE));
				c->launches++;
			}
			{
				Timed t_(c, CLS_FFT_A2A);
				if (!overlap) { k_push_bwd<<<pg, tb, 0, xs>>>(A, X, N, nh, c->nzl, c->nkyl, c->rank, 0, nc, 0, c->nkyl); KERNEL_CHECK(c); }
				GEVB_TRY(rank_barrier(c, xs));
				if (overlap) { CUDA_TRY(cudaEventRecord(c->xev[GEVB_NXEV], xs)); CUDA_TRY(cudaStreamWaitEvent(c->stream, c->xev[GEVB_NXEV], 0)); }
			}
			for (int k = 0; k < nc; k++)
				CUFFT_TRY(cufftExecZ2D(p->bwd2d, (cufftDoubleComplex *) (Xl + (size_t) k * comp_sites), rbulk + k * rf->comp_stride));
			c->launches += nc;
		}
		return 0;
	}
	if (direction == GEVB_FFT_FORWARD)
	{
		// A: [c][ky][zl][kx] -- the ky-slab of peer r is the contiguous chunk r of component c
		for (int k = 0; k < nc; k++)
			CUFFT_TRY(cufftExecD2Z(p->fwd2d, rbulk + k * rf->comp_stride, (cufftDoubleComplex *) (A + (size_t) k * comp_sites)));
		c->launches += nc;
		{ Timed t_(c, CLS_FFT_A2A); GEVB_TRY(alltoall(c, A, B, chunk, nc, false, true)); }   // B: [src][c][kyl][zl of src][kx]
		{
			Timed t_(c, CLS_FFT_TRANSPOSE);
			k_unpack_fwd<<<tg, tb, 0, c->stream>>>(B, (double2 *) cf->data, N, nh, c->nzl, c->nkyl, nc, c->nranks);
			KERNEL_CHECK(c);
		}
		CUFFT_TRY(cufftExecZ2Z(p->z1d, (cufftDoubleComplex *) cf->data, (cufftDoubleComplex *) cf->data, CUFFT_FORWARD));
		c->launches++;
	}
	else
	{
		CUFFT_TRY(cufftExecZ2Z(p->z1d, (cufftDoubleComplex *) cf->data, (cufftDoubleComplex *) A, CUFFT_INVERSE));
		c->launches++;
		{
			Timed t_(c, CLS_FFT_TRANSPOSE);
			k_pack_bwd<<<tg, tb, 0, c->stream>>>(A, B, N, nh, c->nzl, c->nkyl, nc, c->nranks);  // B: [dst][c][kyl][zl of dst][kx]
			KERNEL_CHECK(c);
		}
		{ Timed t_(c, CLS_FFT_A2A); GEVB_TRY(alltoall(c, B, A, chunk, nc, true, false)); }   // A: [c][src][kyl][zl][kx] == [c][ky][zl][kx]
		for (int k = 0; k < nc; k++)
			CUFFT_TRY(cufftExecZ2D(p->bwd2d, (cufftDoubleComplex *) (A + (size_t) k * comp_sites), rbulk + k * rf->comp_stride));
		c->launches += nc;
	}
	return 0;
}
