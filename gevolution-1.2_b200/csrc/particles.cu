// particles.cu -- brick-major cell-sorted FP64 structure-of-arrays particle storage
//
// Replaces LATfield2's Particles<part_simple,...> container (per-cell
// std::list<part_simple>, reference uses at gevolution.hpp:960,977 and
// ic_basic.hpp:1429,1990) by seven flat device arrays {x,y,z,qx,qy,qz,id}
// kept sorted by the key (brick << 9 | cell in brick), cell = floor(pos/dx).
// The integer contract (which cell a particle is filed under, particles per
// cell) is bit-exact with the reference's floor(pos/dx) filing rule.
//
// Re-filing after a drift (the list splice inside LATfield2's moveParticles) is a
// counting sort: the drift kernel accumulates the histogram of the new keys, one
// exclusive scan gives cell_start[], one pass moves the 56-byte records to their
// slots.  Particles move by a fraction of a cell per step, so source and
// destination order are almost the same and both sides of the move stay coalesced.
#include <cub/device/device_scan.cuh>
#include <thread>
#include "gevb_internal.cuh"

namespace {

// host AoS chunk -> SoA append, keeping only particles filed in this rank's slab
__global__ void k_append(int64_t n, const int64_t * __restrict__ id, const double * __restrict__ pos, const double * __restrict__ vel,
                         int N, int z0, int nzl, double dx,
                         double * x, double * y, double * z, double * qx, double * qy, double * qz, int64_t * oid,
                         unsigned long long * counter, int64_t cap)
{
	for (int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x)
	{
		const double pz = pos[3 * i + 2];
		unsigned long long slot;
		if (counter == NULL) slot = (unsigned long long) (cap + i);     // single rank: everything is local, cap = first free slot
		else
		{
			const int cz = cell_of(pz, dx, N);
			if (cz < z0 || cz >= z0 + nzl) continue;
			slot = atomicAdd(counter, 1ull);
			if ((int64_t) slot >= cap) continue;
		}
		x[slot] = pos[3 * i]; y[slot] = pos[3 * i + 1]; z[slot] = pz;
		qx[slot] = vel[3 * i]; qy[slot] = vel[3 * i + 1]; qz[slot] = vel[3 * i + 2];
		oid[slot] = id[i];
	}
}

__global__ void k_count_local(int64_t n, const double * __restrict__ pos, int N, int z0, int nzl, double dx, unsigned long long * counter)
{
	unsigned long long mine = 0;
	for (int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x)
	{
		const int cz = cell_of(pos[3 * i + 2], dx, N);
		mine += (cz >= z0 && cz < z0 + nzl);
	}
	for (int o = 16; o > 0; o >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, o);
	if ((threadIdx.x & 31) == 0 && mine) atomicAdd(counter, mine);
}

#define SCATTER_ILP 4
// keys of particles [first, first + n) from their positions, accumulated into the histogram
__global__ void k_make_keys(BrickGeom G, int64_t first, int64_t n, const double * __restrict__ x, const double * __restrict__ y, const double * __restrict__ z,
                            double dx, uint32_t * __restrict__ key, uint32_t * count, uint32_t * __restrict__ rank)
{
	for (int64_t i = first + blockIdx.x * (int64_t) blockDim.x + threadIdx.x; i < first + n; i += (int64_t) gridDim.x * blockDim.x)
	{
		const int cx = cell_of(x[i], dx, G.N), cy = cell_of(y[i], dx, G.N), cz = cell_of(z[i], dx, G.N) - G.z0;
		const uint32_t k = brick_key(G, cx, cy, cz);
		key[i] = k;
		const uint32_t r = atomicAdd(count + k, 1u);
		if (rank) rank[i] = r;
	}
}

// the move of the counting sort when every particle knows its rank inside its new cell (rebin_variant = 1): no atomic,
// two independent streaming loads (key, rank), one dependent lookup of cell_start (L2: neighbours share lines), seven
// independent record loads, seven stores
__global__ void __launch_bounds__(256) k_scatter_ranked(int64_t n, const unsigned long long * __restrict__ n_dev, const uint32_t * __restrict__ key, const uint32_t * __restrict__ rank, const uint32_t * __restrict__ cell_start,
                          const double * __restrict__ x, const double * __restrict__ y, const double * __restrict__ z,
                          const double * __restrict__ qx, const double * __restrict__ qy, const double * __restrict__ qz, const int64_t * __restrict__ id,
                          double * __restrict__ ox, double * __restrict__ oy, double * __restrict__ oz,
                          double * __restrict__ oqx, double * __restrict__ oqy, double * __restrict__ oqz, int64_t * __restrict__ oid)
{
	const int64_t stride = (int64_t) gridDim.x * blockDim.x;
	if (n_dev) n = (int64_t) *n_dev;                            // after a migration over peer memory the count lives on the device
	for (int64_t i0 = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; i0 < n; i0 += SCATTER_ILP * stride)
	{
		uint32_t k[SCATTER_ILP], d[SCATTER_ILP];
		double v[SCATTER_ILP][6];
		int64_t vid[SCATTER_ILP];
		#pragma unroll
		for (int u = 0; u < SCATTER_ILP; u++)
		{
			const int64_t i = i0 + u * stride;
			k[u] = GEVB_INVALID_KEY; d[u] = 0;
			if (i < n) { k[u] = __ldcs(key + i); d[u] = __ldcs(rank + i); }
		}
		#pragma unroll
		for (int u = 0; u < SCATTER_ILP; u++)
		{
			const int64_t i = i0 + u * stride;
			if (k[u] == GEVB_INVALID_KEY) continue;
			d[u] += __ldg(cell_start + k[u]);
			v[u][0] = __ldcs(x + i); v[u][1] = __ldcs(y + i); v[u][2] = __ldcs(z + i);
			v[u][3] = __ldcs(qx + i); v[u][4] = __ldcs(qy + i); v[u][5] = __ldcs(qz + i);
			vid[u] = __ldcs(id + i);
		}
		#pragma unroll
		for (int u = 0; u < SCATTER_ILP; u++)
		{
			if (k[u] == GEVB_INVALID_KEY) continue;
			ox[d[u]] = v[u][0]; oy[d[u]] = v[u][1]; oz[d[u]] = v[u][2];
			oqx[d[u]] = v[u][3]; oqy[d[u]] = v[u][4]; oqz[d[u]] = v[u][5];
			oid[d[u]] = vid[u];
		}
	}
}

// the move of the counting sort through a shared-memory window (rebin_variant = 2, 3).  The records of a chunk of
// WIN_CHUNK consecutive particles mostly stay where they are in the sorted order (a particle that keeps its brick moves
// by at most a few hundred slots), so their new slots fall into a window of WIN_SLOTS consecutive slots around the slot of the
// chunk's first particle.  The block drops those records into the window in shared memory and writes the window out row by
// row: consecutive lanes store consecutive slots (whole 128-byte lines but for the few slots other blocks fill) instead of
// 8-byte stores scattered over a dozen sectors per request.  Records that leave the window are stored directly, as in
// k_scatter.  Where the window lies only matters for speed.  RANKED: the slot inside the cell comes from the rank the drift
// kernel recorded (rebin_variant 3), else from counting the histogram down.
#define WIN_PAD 128
#define WIN_THREADS 256
#define WIN_SMEM(chunk) (7 * ((chunk) + 2 * WIN_PAD) * 8 + ((chunk) + 2 * WIN_PAD))
template <bool RANKED, int WIN_CHUNK>
__global__ void __launch_bounds__(WIN_THREADS, WIN_CHUNK == 1024 ? 3 : 5) k_scatter_window(int64_t n, const unsigned long long * __restrict__ n_dev, const uint32_t * __restrict__ key, const uint32_t * __restrict__ rank,
                          const uint32_t * __restrict__ cell_start, uint32_t * count,
                          const double * __restrict__ x, const double * __restrict__ y, const double * __restrict__ z,
                          const double * __restrict__ qx, const double * __restrict__ qy, const double * __restrict__ qz, const int64_t * __restrict__ id,
                          double * __restrict__ ox, double * __restrict__ oy, double * __restrict__ oz,
                          double * __restrict__ oqx, double * __restrict__ oqy, double * __restrict__ oqz, int64_t * __restrict__ oid)
{
	constexpr int WIN_SLOTS = WIN_CHUNK + 2 * WIN_PAD, WIN_PER = WIN_CHUNK / WIN_THREADS;
	extern __shared__ __align__(16) unsigned char win_smem[];
	double * buf = (double *) win_smem;                                   // [7][WIN_SLOTS]
	unsigned char * filled = win_smem + 7 * WIN_SLOTS * sizeof(double);   // [WIN_SLOTS]
	__shared__ uint32_t base_s;
	if (n_dev) n = (int64_t) *n_dev;
	const int64_t nchunks = (n + WIN_CHUNK - 1) / WIN_CHUNK;
	uint32_t k[WIN_PER], d[WIN_PER];
	double v[WIN_PER][6];
	int64_t vid[WIN_PER];
	// the records, keys and slots of a chunk: issued one chunk ahead, so that they are in flight while the previous window is written out
	auto fetch = [&](int64_t ch)
	{
		const int64_t s0 = ch * WIN_CHUNK;
		#pragma unroll
		for (int u = 0; u < WIN_PER; u++)
		{
			const int64_t i = s0 + u * WIN_THREADS + threadIdx.x;
			k[u] = GEVB_INVALID_KEY; d[u] = 0;
			if (ch < nchunks && i < n) { k[u] = __ldcs(key + i); if (RANKED) d[u] = __ldcs(rank + i); }
		}
		#pragma unroll
		for (int u = 0; u < WIN_PER; u++)
		{
			const int64_t i = s0 + u * WIN_THREADS + threadIdx.x;
			if (k[u] == GEVB_INVALID_KEY) continue;
			v[u][0] = __ldcs(x + i); v[u][1] = __ldcs(y + i); v[u][2] = __ldcs(z + i);
			v[u][3] = __ldcs(qx + i); v[u][4] = __ldcs(qy + i); v[u][5] = __ldcs(qz + i);
			vid[u] = __ldcs(id + i);
		}
		#pragma unroll
		for (int u = 0; u < WIN_PER; u++)
			if (k[u] != GEVB_INVALID_KEY) d[u] = RANKED ? d[u] + __ldg(cell_start + k[u]) : __ldg(cell_start + k[u]) + atomicSub(count + k[u], 1u) - 1u;
	};
	fetch(blockIdx.x);
	for (int64_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x)
	{
		for (int w = threadIdx.x; w < WIN_SLOTS / 4; w += WIN_THREADS) ((uint32_t *) filled)[w] = 0u;
		if (threadIdx.x == 0) base_s = k[0] != GEVB_INVALID_KEY ? (d[0] > WIN_PAD ? d[0] - WIN_PAD : 0u) : 0xffffffffu - WIN_SLOTS;
		__syncthreads();
		const uint32_t base = base_s;
		#pragma unroll
		for (int u = 0; u < WIN_PER; u++)
		{
			if (k[u] == GEVB_INVALID_KEY) continue;
			const uint32_t off = d[u] - base;                               // unsigned: slots below the window wrap to huge values
			if (off < WIN_SLOTS)
			{
				#pragma unroll
				for (int a = 0; a < 6; a++) buf[a * WIN_SLOTS + off] = v[u][a];
				((int64_t *) buf)[6 * WIN_SLOTS + off] = vid[u];
				filled[off] = 1;
			}
			else
			{
				ox[d[u]] = v[u][0]; oy[d[u]] = v[u][1]; oz[d[u]] = v[u][2];
				oqx[d[u]] = v[u][3]; oqy[d[u]] = v[u][4]; oqz[d[u]] = v[u][5];
				oid[d[u]] = vid[u];
			}
		}
		__syncthreads();
		fetch(ch + gridDim.x);
		for (int j = threadIdx.x; j < WIN_SLOTS; j += WIN_THREADS)
		{
			if (!filled[j]) continue;
			const size_t t = (size_t) base + j;
			ox[t] = buf[j]; oy[t] = buf[WIN_SLOTS + j]; oz[t] = buf[2 * WIN_SLOTS + j];
			oqx[t] = buf[3 * WIN_SLOTS + j]; oqy[t] = buf[4 * WIN_SLOTS + j]; oqz[t] = buf[5 * WIN_SLOTS + j];
			oid[t] = ((const int64_t *) buf)[6 * WIN_SLOTS + j];
		}
		__syncthreads();
	}
}

static unsigned window_grid(gevb_ctx * c, int64_t n, int chunk)
{
	const int64_t chunks = (n + chunk - 1) / chunk, persistent = (int64_t) c->num_sms * (chunk == 1024 ? 3 : 5);
	return (unsigned) (chunks < persistent ? (chunks > 0 ? chunks : 1) : persistent);
}

template <bool RANKED, int CHUNK>
static int launch_window(gevb_ctx * c, int64_t n_in, gevb_pcls * p, int s, int d)
{
	CUDA_TRY(cudaFuncSetAttribute(k_scatter_window<RANKED, CHUNK>, cudaFuncAttributeMaxDynamicSharedMemorySize, WIN_SMEM(CHUNK)));
	k_scatter_window<RANKED, CHUNK><<<window_grid(c, n_in, CHUNK), WIN_THREADS, WIN_SMEM(CHUNK), c->stream>>>(n_in, p->d_nin, p->key, RANKED ? p->rank : NULL, p->cell_start, p->cell_count,
		p->x[s], p->y[s], p->z[s], p->qx[s], p->qy[s], p->qz[s], p->id[s],
		p->x[d], p->y[d], p->z[d], p->qx[d], p->qy[d], p->qz[d], p->id[d]);
	KERNEL_CHECK(c);
	return 0;
}

// the move of the counting sort: slot = cell_start[key] + (number of particles of this cell not yet placed) - 1.
// Counting down leaves cell_count all zero again, ready for the next histogram.  Four particles per thread and
// iteration, all loads and atomics of the four issued before the first dependent store (the chain key -> atomic ->
// slot -> stores is latency bound otherwise).
__global__ void __launch_bounds__(256) k_scatter(int64_t n, const unsigned long long * __restrict__ n_dev, const uint32_t * __restrict__ key, const uint32_t * __restrict__ cell_start, uint32_t * count,
                          const double * __restrict__ x, const double * __restrict__ y, const double * __restrict__ z,
                          const double * __restrict__ qx, const double * __restrict__ qy, const double * __restrict__ qz, const int64_t * __restrict__ id,
                          double * __restrict__ ox, double * __restrict__ oy, double * __restrict__ oz,
                          double * __restrict__ oqx, double * __restrict__ oqy, double * __restrict__ oqz, int64_t * __restrict__ oid)
{
	const int64_t stride = (int64_t) gridDim.x * blockDim.x;
	if (n_dev) n = (int64_t) *n_dev;
	for (int64_t i0 = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; i0 < n; i0 += SCATTER_ILP * stride)
	{
		uint32_t k[SCATTER_ILP], d[SCATTER_ILP];
		#pragma unroll
		for (int u = 0; u < SCATTER_ILP; u++)
		{
			const int64_t i = i0 + u * stride;
			k[u] = i < n ? __ldcs(key + i) : GEVB_INVALID_KEY;    // GEVB_INVALID_KEY: left the slab (sent to a neighbour rank)
		}
		#pragma unroll
		for (int u = 0; u < SCATTER_ILP; u++)
			if (k[u] != GEVB_INVALID_KEY) d[u] = __ldg(cell_start + k[u]) + atomicSub(count + k[u], 1u) - 1u;
		#pragma unroll
		for (int u = 0; u < SCATTER_ILP; u++)
		{
			const int64_t i = i0 + u * stride;
			if (k[u] == GEVB_INVALID_KEY) continue;
			const double vx = __ldcs(x + i), vy = __ldcs(y + i), vz = __ldcs(z + i), wx = __ldcs(qx + i), wy = __ldcs(qy + i), wz = __ldcs(qz + i);
			const int64_t vid = __ldcs(id + i);
			ox[d[u]] = vx; oy[d[u]] = vy; oz[d[u]] = vz;
			oqx[d[u]] = wx; oqy[d[u]] = wy; oqz[d[u]] = wz;
			oid[d[u]] = vid;
		}
	}
}

// per-cell counts in lattice order [zl][y][x] from the prefix sums in key order
__global__ void k_counts_out(BrickGeom G, const uint32_t * __restrict__ cell_start, uint32_t * __restrict__ counts)
{
	const size_t cells = (size_t) G.nzl * G.N * G.N;
	for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < cells; i += (size_t) gridDim.x * blockDim.x)
	{
		const int cx = (int) (i % G.N); const size_t r = i / G.N;
		const int cy = (int) (r % G.N), cz = (int) (r / G.N);
		const uint32_t k = brick_key(G, cx, cy, cz);
		counts[i] = cell_start[k + 1] - cell_start[k];
	}
}

__global__ void k_interleave3(int64_t n, const double * __restrict__ a, const double * __restrict__ b, const double * __restrict__ c, double * __restrict__ out)
{
	for (int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x)
	{
		out[3 * i] = a[i]; out[3 * i + 1] = b[i]; out[3 * i + 2] = c[i];
	}
}

void free_arrays(gevb_pcls * p)
{
	for (int b = 0; b < 2; b++)
	{
		cudaFree(p->x[b]); cudaFree(p->y[b]); cudaFree(p->z[b]); cudaFree(p->qx[b]); cudaFree(p->qy[b]); cudaFree(p->qz[b]);
		cudaFree(p->id[b]);
		p->x[b] = p->y[b] = p->z[b] = p->qx[b] = p->qy[b] = p->qz[b] = NULL; p->id[b] = NULL;
	}
	cudaFree(p->key); p->key = NULL;
	cudaFree(p->rank); p->rank = NULL;
}

} // namespace

extern "C" int gevb_pcls_create(gevb_ctx * c, gevb_pcls ** out, double mass)
{
	GEVB_CHECK_ARG(c != NULL && out != NULL, "gevb_pcls_create: NULL argument");
	BrickGeom G;
	G.N = c->N; G.nzl = c->nzl; G.z0 = c->z0;
	const int spanx = GEVB_BX << GEVB_SX_BITS, spany = GEVB_BY << GEVB_SY_BITS, spanz = GEVB_BZ << GEVB_SZ_BITS;   // cells per super-brick edge
	G.nsx = (c->N + spanx - 1) / spanx; G.nsy = (c->N + spany - 1) / spany; G.nsz = (c->nzl + spanz - 1) / spanz;
	G.sx_shift = G.sy_shift = -1;
	if ((G.nsx & (G.nsx - 1)) == 0 && (G.nsy & (G.nsy - 1)) == 0)
	{
		G.sx_shift = 0; while ((1 << G.sx_shift) < G.nsx) G.sx_shift++;
		G.sy_shift = 0; while ((1 << G.sy_shift) < G.nsy) G.sy_shift++;
	}
	const uint64_t ncells = ((uint64_t) G.nsx * G.nsy * G.nsz << GEVB_SUPER_BITS) * GEVB_BRICK_CELLS;
	GEVB_CHECK_ARG(ncells < (1ull << 31), "gevb_pcls_create: local slab has more than 2^31 cells");
	G.nbricks = (uint32_t) (ncells / GEVB_BRICK_CELLS); G.ncells = (uint32_t) ncells;
	CUDA_TRY(cudaSetDevice(c->device));
	gevb_pcls * p = new gevb_pcls();
	memset(p, 0, sizeof(*p));
	p->ctx = c; p->mass = mass; p->geom = G;
	const size_t bytes = ((size_t) G.ncells + 1) * sizeof(uint32_t);
	cudaError_t e1 = cudaMalloc(&p->cell_count, bytes), e2 = cudaMalloc(&p->cell_start, bytes);
	if (e1 != cudaSuccess || e2 != cudaSuccess)
	{
		cudaFree(p->cell_count); cudaFree(p->cell_start); delete p;
		GEVB_FAIL("gevb_pcls_create: cudaMalloc of the cell tables (2 x %zu bytes) failed", bytes);
	}
	CUDA_TRY(cudaMemsetAsync(p->cell_count, 0, bytes, c->stream));
	CUDA_TRY(cudaMemsetAsync(p->cell_start, 0, bytes, c->stream));
	*out = p;
	return 0;
}

extern "C" int gevb_pcls_destroy(gevb_pcls * p)
{
	if (p == NULL) return 0;
	cudaSetDevice(p->ctx->device);
	cudaStreamSynchronize(p->ctx->stream);
	free_arrays(p);
	cudaFree(p->cell_count); cudaFree(p->cell_start);
	delete p;
	return 0;
}

extern "C" double gevb_pcls_mass(gevb_pcls * p) { return p ? p->mass : 0.; }

extern "C" void gevb_brick_dims(int * brick3, int * super3)
{
	if (brick3) { brick3[0] = GEVB_BX; brick3[1] = GEVB_BY; brick3[2] = GEVB_BZ; }
	if (super3) { super3[0] = 1 << GEVB_SX_BITS; super3[1] = 1 << GEVB_SY_BITS; super3[2] = 1 << GEVB_SZ_BITS; }
}

// grow capacity, keeping the live particles (and their keys)
int gevb_pcls_reserve(gevb_pcls * p, int64_t cap)
{
	if (cap <= p->cap) return 0;
	gevb_ctx * c = p->ctx;
	GEVB_CHECK_ARG(cap < (1ll << 32) - 1, "particles: more than 2^32 particles on one rank");
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	gevb_pcls old = *p;
	for (int b = 0; b < 2; b++)
	{
		double ** arrs[6] = {&p->x[b], &p->y[b], &p->z[b], &p->qx[b], &p->qy[b], &p->qz[b]};
		for (int a = 0; a < 6; a++) CUDA_TRY(cudaMalloc(arrs[a], sizeof(double) * (cap + 2)));   // + 2: the deposit's bulk copies round a range out to 16-byte boundaries
		CUDA_TRY(cudaMalloc(&p->id[b], sizeof(int64_t) * cap));
	}
	CUDA_TRY(cudaMalloc(&p->key, sizeof(uint32_t) * cap));
	CUDA_TRY(cudaMalloc(&p->rank, sizeof(uint32_t) * cap));
	if (old.cap > 0)
	{
		// the live arrays may hold more than n records (received particles appended behind them)
		const int s = old.cur, d = 0;
		const size_t m = (size_t) old.cap;
		CUDA_TRY(cudaMemcpy(p->x[d], old.x[s], sizeof(double) * m, cudaMemcpyDeviceToDevice));
		CUDA_TRY(cudaMemcpy(p->y[d], old.y[s], sizeof(double) * m, cudaMemcpyDeviceToDevice));
		CUDA_TRY(cudaMemcpy(p->z[d], old.z[s], sizeof(double) * m, cudaMemcpyDeviceToDevice));
		CUDA_TRY(cudaMemcpy(p->qx[d], old.qx[s], sizeof(double) * m, cudaMemcpyDeviceToDevice));
		CUDA_TRY(cudaMemcpy(p->qy[d], old.qy[s], sizeof(double) * m, cudaMemcpyDeviceToDevice));
		CUDA_TRY(cudaMemcpy(p->qz[d], old.qz[s], sizeof(double) * m, cudaMemcpyDeviceToDevice));
		CUDA_TRY(cudaMemcpy(p->id[d], old.id[s], sizeof(int64_t) * m, cudaMemcpyDeviceToDevice));
		CUDA_TRY(cudaMemcpy(p->key, old.key, sizeof(uint32_t) * m, cudaMemcpyDeviceToDevice));
	}
	p->cur = 0; p->cap = cap;
	free_arrays(&old);
	return 0;
}

// restore the brick-major cell order: [histogram ->] exclusive scan -> scatter
int gevb_pcls_rebin(gevb_pcls * p, int64_t n_in, int64_t n_out, bool hist_valid)
{
	gevb_ctx * c = p->ctx;
	const BrickGeom & G = p->geom;
	const int s = p->cur, d = 1 - p->cur;
	const double dx = 1.0 / (double) c->N;
	// the histogram's producers (here, the drift kernel, the append of received particles) all follow the knob, so a
	// histogram passed in as valid comes with ranks exactly when the knob is on; it must not change between a drift and its re-bin
	const int rebin_variant = gevb_tune(TUNE_REBIN_VARIANT);
	const bool ranked = (rebin_variant & 1) != 0, windowed = (rebin_variant & 2) != 0, small_window = (rebin_variant & 4) != 0;
	if (!hist_valid && n_in > 0)
	{
		k_make_keys<<<gevb_grid(c, (size_t) n_in, 256), 256, 0, c->stream>>>(G, 0, n_in, p->x[s], p->y[s], p->z[s], dx, p->key, p->cell_count, ranked ? p->rank : NULL);
		KERNEL_CHECK(c);
	}
	size_t temp_bytes = 0;
	const int items = (int) (G.ncells + 1);
	CUDA_TRY(cub::DeviceScan::ExclusiveSum(NULL, temp_bytes, p->cell_count, p->cell_start, items, c->stream));
	void * temp;
	GEVB_TRY(gevb_ctx_scratch(c, temp_bytes, &temp));
	CUDA_TRY(cub::DeviceScan::ExclusiveSum(temp, temp_bytes, p->cell_count, p->cell_start, items, c->stream));
	c->launches += 2;
	if (ranked)
	{
		// nothing counts the histogram down: clear it for the next one
		CUDA_TRY(cudaMemsetAsync(p->cell_count, 0, ((size_t) G.ncells + 1) * sizeof(uint32_t), c->stream));
		if (n_in > 0 && windowed)
		{
			GEVB_TRY((small_window ? launch_window<true, 512> : launch_window<true, 1024>)(c, n_in, p, s, d));
		}
		else if (n_in > 0)
		{
			k_scatter_ranked<<<gevb_grid(c, (size_t) n_in, 256), 256, 0, c->stream>>>(n_in, p->d_nin, p->key, p->rank, p->cell_start,
				p->x[s], p->y[s], p->z[s], p->qx[s], p->qy[s], p->qz[s], p->id[s],
				p->x[d], p->y[d], p->z[d], p->qx[d], p->qy[d], p->qz[d], p->id[d]);
			KERNEL_CHECK(c);
		}
	}
	else if (n_in > 0 && windowed)
	{
		GEVB_TRY((small_window ? launch_window<false, 512> : launch_window<false, 1024>)(c, n_in, p, s, d));
	}
	else if (n_in > 0)
	{
		k_scatter<<<gevb_grid(c, (size_t) n_in, 256), 256, 0, c->stream>>>(n_in, p->d_nin, p->key, p->cell_start, p->cell_count,
			p->x[s], p->y[s], p->z[s], p->qx[s], p->qy[s], p->qz[s], p->id[s],
			p->x[d], p->y[d], p->z[d], p->qx[d], p->qy[d], p->qz[d], p->id[d]);
		KERNEL_CHECK(c);
	}
	p->cur = d;
	p->n = n_out;
	return 0;
}

extern "C" int gevb_pcls_add(gevb_pcls * p, int64_t n, const int64_t * id, const double * pos, const double * vel)
{
	GEVB_CHECK_ARG(p != NULL && n >= 0, "gevb_pcls_add: bad arguments");
	if (n == 0) return 0;
	GEVB_CHECK_ARG(id != NULL && pos != NULL && vel != NULL, "gevb_pcls_add: NULL array");
	gevb_ctx * c = p->ctx;
	CUDA_TRY(cudaSetDevice(c->device));
	const double dx = 1.0 / (double) c->N;
	// the host arrays are staged in chunks of at most 2^26 particles (3.5 GB) and transposed into the SoA arrays;
	// no host synchronisation inside the loop on a single rank, so the copies run back to back at PCIe speed
	const int64_t chunk = (int64_t) 1 << 26;
	const int64_t n_before = p->n;
	unsigned long long * counter = (unsigned long long *) (c->d_red + 4000);
	if (c->nranks == 1) GEVB_TRY(gevb_pcls_reserve(p, p->n + n));
	int64_t nloc_host = 0;
	if (c->nranks > 1)
	{
		// which of the offered particles are filed in this rank's slab is decided by the same floor(pos/dx) the device uses; counting
		// them on the host first (a few threads over the z coordinates) sizes the arrays once, so that the upload below runs
		// chunk after chunk at PCIe speed without a count-and-synchronise round trip per chunk
		// When the arrays already hold all n offered particles (a refill of the same container, as in a restart or an
		// end-to-end step), nothing needs sizing and the count is skipped.
		const bool roomy = p->cap >= p->n + n;
		const int nthreads = roomy ? 0 : n > (1 << 20) ? 4 : 1;
		std::vector<int64_t> part(nthreads > 0 ? nthreads : 1, 0);
		auto count = [&](int t)
		{
			int64_t mine = 0;
			for (int64_t i = n * t / nthreads; i < n * (t + 1) / nthreads; i++)
			{
				int cz = (int) floor(pos[3 * i + 2] / dx);
				cz = cz >= c->N ? c->N - 1 : (cz < 0 ? 0 : cz);
				mine += (cz >= c->z0 && cz < c->z0 + c->nzl);
			}
			part[t] = mine;
		};
		if (nthreads == 1) count(0);
		else if (nthreads > 1)
		{
			std::vector<std::thread> pool;
			for (int t = 0; t < nthreads; t++) pool.emplace_back(count, t);
			for (std::thread & th : pool) th.join();
		}
		for (int t = 0; t < nthreads; t++) nloc_host += part[t];
		if (roomy) nloc_host = -1;                                                    // not counted
		else
		{
			if (nloc_host == 0) return 0;
			GEVB_TRY(gevb_pcls_reserve(p, p->n + nloc_host));
		}
		c->h_red[8] = 0.;
		unsigned long long start = (unsigned long long) p->n;
		memcpy(c->h_red + 8, &start, sizeof(start));                                  // pinned: stays valid while the copy is in flight
		CUDA_TRY(cudaMemcpyAsync(counter, c->h_red + 8, sizeof(start), cudaMemcpyHostToDevice, c->stream));
	}
	for (int64_t off = 0; off < n; off += chunk)
	{
		const int64_t m = (n - off < chunk) ? n - off : chunk;
		void * stage;
		GEVB_TRY(gevb_ctx_scratch(c, (size_t) m * 56, &stage));
		int64_t * did = (int64_t *) stage;
		double * dpos = (double *) (did + m), * dvel = dpos + 3 * m;
		CUDA_TRY(cudaMemcpyAsync(did, id + off, sizeof(int64_t) * m, cudaMemcpyHostToDevice, c->stream));
		CUDA_TRY(cudaMemcpyAsync(dpos, pos + 3 * off, sizeof(double) * 3 * m, cudaMemcpyHostToDevice, c->stream));
		CUDA_TRY(cudaMemcpyAsync(dvel, vel + 3 * off, sizeof(double) * 3 * m, cudaMemcpyHostToDevice, c->stream));
		if (c->nranks == 1)
		{
			const int b = p->cur;
			k_append<<<gevb_grid(c, (size_t) m, 256), 256, 0, c->stream>>>(m, did, dpos, dvel, c->N, c->z0, c->nzl, dx,
				p->x[b], p->y[b], p->z[b], p->qx[b], p->qy[b], p->qz[b], p->id[b], NULL, p->n);
			KERNEL_CHECK(c);
			p->n += m;
			continue;
		}
		// several ranks: the chunk's local particles are appended behind the ones kept so far (the device counter runs on from
		// chunk to chunk; room for all of them was reserved from the host-side count above), no host synchronisation in the loop
		const int b = p->cur;
		k_append<<<gevb_grid(c, (size_t) m, 256), 256, 0, c->stream>>>(m, did, dpos, dvel, c->N, c->z0, c->nzl, dx,
			p->x[b], p->y[b], p->z[b], p->qx[b], p->qy[b], p->qz[b], p->id[b], counter, p->cap);
		KERNEL_CHECK(c);
	}
	if (c->nranks > 1)
	{
		unsigned long long total = 0;
		CUDA_TRY(cudaMemcpyAsync(&total, counter, sizeof(total), cudaMemcpyDeviceToHost, c->stream));
		CUDA_TRY(cudaStreamSynchronize(c->stream));
		GEVB_CHECK_ARG(nloc_host < 0 ? (int64_t) total <= p->cap : (int64_t) total == n_before + nloc_host,
			"gevb_pcls_add: the device kept %llu particles, the host counted %lld (capacity %lld)", total, (long long) (n_before + nloc_host), (long long) p->cap);
		p->n = (int64_t) total;
	}
	if (p->n == n_before) return 0;
	// keys + histogram of everything (old particles included: their keys are not kept between calls)
	return gevb_pcls_rebin(p, p->n, p->n, false);
}

// Particles::initialize on an existing container: drops the particles, keeps the device arrays
extern "C" int gevb_pcls_reset(gevb_pcls * p, double mass)
{
	GEVB_CHECK_ARG(p != NULL, "gevb_pcls_reset: NULL handle");
	gevb_ctx * c = p->ctx;
	CUDA_TRY(cudaSetDevice(c->device));
	p->n = 0; p->mass = mass;
	CUDA_TRY(cudaMemsetAsync(p->cell_start, 0, ((size_t) p->geom.ncells + 1) * sizeof(uint32_t), c->stream));
	return 0;
}

extern "C" int gevb_pcls_count(gevb_pcls * p, int64_t * n_local)
{
	GEVB_CHECK_ARG(p != NULL && n_local != NULL, "gevb_pcls_count: NULL argument");
	*n_local = p->n;
	return 0;
}

extern "C" int gevb_pcls_download(gevb_pcls * p, int64_t * id, double * pos, double * vel)
{
	GEVB_CHECK_ARG(p != NULL, "gevb_pcls_download: NULL handle");
	gevb_ctx * c = p->ctx;
	if (p->n == 0) return 0;
	CUDA_TRY(cudaSetDevice(c->device));
	const int b = p->cur;
	void * stage;
	GEVB_TRY(gevb_ctx_scratch(c, (size_t) p->n * 48, &stage));
	double * spos = (double *) stage, * svel = spos + 3 * p->n;
	const int grid = gevb_grid(c, (size_t) p->n, 256);
	// all device work is queued before the first copy so that the three copies run back to back
	if (pos) { k_interleave3<<<grid, 256, 0, c->stream>>>(p->n, p->x[b], p->y[b], p->z[b], spos); KERNEL_CHECK(c); }
	if (vel) { k_interleave3<<<grid, 256, 0, c->stream>>>(p->n, p->qx[b], p->qy[b], p->qz[b], svel); KERNEL_CHECK(c); }
	if (id) CUDA_TRY(cudaMemcpyAsync(id, p->id[b], sizeof(int64_t) * p->n, cudaMemcpyDeviceToHost, c->stream));
	if (pos) CUDA_TRY(cudaMemcpyAsync(pos, spos, sizeof(double) * 3 * p->n, cudaMemcpyDeviceToHost, c->stream));
	if (vel) CUDA_TRY(cudaMemcpyAsync(vel, svel, sizeof(double) * 3 * p->n, cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	return 0;
}

extern "C" int gevb_pcls_cell_counts(gevb_pcls * p, uint32_t * counts)
{
	GEVB_CHECK_ARG(p != NULL && counts != NULL, "gevb_pcls_cell_counts: NULL argument");
	gevb_ctx * c = p->ctx;
	CUDA_TRY(cudaSetDevice(c->device));
	const size_t cells = (size_t) c->nzl * c->plane();
	void * stage;
	GEVB_TRY(gevb_ctx_scratch(c, cells * sizeof(uint32_t), &stage));
	k_counts_out<<<gevb_grid(c, cells, 256), 256, 0, c->stream>>>(p->geom, p->cell_start, (uint32_t *) stage);
	KERNEL_CHECK(c);
	CUDA_TRY(cudaMemcpyAsync(counts, stage, cells * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	return 0;
}
