// ctx.cu -- context, field storage, ghost-plane exchange, reductions
//
// Replaces, for the hot path only: LATfield2 `parallel`, `Lattice`, `Field<T>`
// storage, Field::updateHalo, projection_init and the three *_comm folds
// (reference call sites main.cpp:152,213-246,378,411,435,450,459-463,518,568,598).
#include <math.h>
#include <stdarg.h>
#include <string.h>
#include "gevb_internal.cuh"

static thread_local char g_err[1024] = "";

void gevb_set_error(const char * fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
}

extern "C" const char * gevb_last_error(void) { return g_err; }
extern "C" const char * gevb_version(void) { return "gevb 0.1 (sm_100a; hot path of gevolution 1.2)"; }

extern "C" int gevb_nccl_unique_id(void * out128)
{
	GEVB_CHECK_ARG(out128 != NULL, "gevb_nccl_unique_id: NULL output");
	ncclUniqueId id;
	static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
	GEVB_TRY(gevb_nccl_load());
	NCCL_TRY(ncclGetUniqueId(&id));
	memcpy(out128, &id, sizeof(id));
	return 0;
}

// ---- tuning knobs: kernel variants kept side by side for ablation runs (bench.py --ablate); the defaults are the
//      measured best.  Environment variables GEVB_<KNOB> (upper case) preset them.
static const char * const tune_names[GEVB_NTUNE] = {"geodesic_variant", "fft_exchange", "fft_overlap", "fft_decomposed", "deposit_variant", "fft_l2_planes", "rebin_variant", "fft_fused", "peer_comm", "geodesic_tma", "tma_l2_promotion", "fft_xpass"};
static int tune_values[GEVB_NTUNE] = {5, 1, 2, 1, 18, 0, 2, 1, 1, 1, 0, 0};
static bool tune_env_read = false;
static void tune_read_env()
{
	if (tune_env_read) return;
	tune_env_read = true;
	for (int k = 0; k < GEVB_NTUNE; k++)
	{
		std::string name = "GEVB_";
		for (const char * q = tune_names[k]; *q; q++) name += (char) toupper(*q);
		const char * e = getenv(name.c_str());
		if (e) tune_values[k] = atoi(e);
	}
}
int gevb_tune(int knob) { tune_read_env(); return tune_values[knob]; }
typedef CUresult (* EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                 CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int gevb_tensor_map_3d(gevb_ctx * c, CUtensorMap * map, const double * base, int box_x, int box_y, int box_z)
{
	static EncodeTiled encode = []() -> EncodeTiled
	{
		void * p = NULL;
		cudaDriverEntryPointQueryResult q;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return NULL;
		return (EncodeTiled) p;
	}();
	GEVB_CHECK_ARG(encode != NULL, "the driver does not provide cuTensorMapEncodeTiled");
	const cuuint64_t dims[3] = {(cuuint64_t) c->N, (cuuint64_t) c->N, (cuuint64_t) c->nzl + 2};
	const cuuint64_t strides[2] = {(cuuint64_t) c->N * sizeof(double), (cuuint64_t) c->N * c->N * sizeof(double)};
	const cuuint32_t box[3] = {(cuuint32_t) box_x, (cuuint32_t) box_y, (cuuint32_t) box_z};
	const cuuint32_t estr[3] = {1, 1, 1};
	static const CUtensorMapL2promotion promo[4] = {CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B};
	const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void *) base, dims, strides, box, estr,
		CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo[tune_values[TUNE_TMA_L2_PROMOTION] & 3], CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	GEVB_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int) r);
	return 0;
}

extern "C" int gevb_tuning(const char * knob, int value)
{
	GEVB_CHECK_ARG(knob != NULL, "gevb_tuning: NULL knob");
	tune_read_env();
	for (int k = 0; k < GEVB_NTUNE; k++)
		if (strcmp(knob, tune_names[k]) == 0) { tune_values[k] = value; return 0; }
	GEVB_FAIL("gevb_tuning: unknown knob '%s'", knob);
}

// the slab decomposition by itself (pure host arithmetic, no device): rank r owns z-planes [z0, z0 + nzl) of real space
// and, after a forward transform, ky-rows [ky0, ky0 + nkyl) of Fourier space
extern "C" int gevb_slab_geometry(int ngrid, int rank, int nranks, int * z0, int * nz_local, int * ky0, int * nky_local)
{
	GEVB_CHECK_ARG(ngrid >= 4 && ngrid % 2 == 0, "gevb_slab_geometry: Ngrid must be even and >= 4 (got %d)", ngrid);
	GEVB_CHECK_ARG(nranks >= 1 && rank >= 0 && rank < nranks, "gevb_slab_geometry: bad rank %d of %d", rank, nranks);
	GEVB_CHECK_ARG(ngrid % nranks == 0, "gevb_slab_geometry: Ngrid %d not divisible by %d ranks", ngrid, nranks);
	GEVB_CHECK_ARG(nranks == 1 || ngrid / nranks >= 2, "gevb_slab_geometry: slabs must be at least 2 planes thick");
	const int n = ngrid / nranks;
	if (z0) *z0 = rank * n;
	if (nz_local) *nz_local = n;
	if (ky0) *ky0 = rank * n;
	if (nky_local) *nky_local = n;
	return 0;
}

extern "C" int gevb_ctx_create(gevb_ctx ** out, int ngrid, int device, int rank, int nranks, const void * nccl_id)
{
	GEVB_CHECK_ARG(out != NULL, "gevb_ctx_create: NULL output");
	GEVB_CHECK_ARG(ngrid >= 4 && ngrid % 2 == 0, "gevb_ctx_create: Ngrid must be even and >= 4 (got %d)", ngrid);
	GEVB_CHECK_ARG(nranks >= 1 && rank >= 0 && rank < nranks, "gevb_ctx_create: bad rank %d of %d", rank, nranks);
	GEVB_CHECK_ARG(ngrid % nranks == 0, "gevb_ctx_create: Ngrid %d not divisible by %d ranks", ngrid, nranks);
	GEVB_CHECK_ARG(nranks == 1 || ngrid / nranks >= 2, "gevb_ctx_create: slabs must be at least 2 planes thick");
	GEVB_CHECK_ARG(nranks <= GEVB_MAX_RANKS, "gevb_ctx_create: at most %d ranks", GEVB_MAX_RANKS);
	GEVB_CHECK_ARG(nranks == 1 || nccl_id != NULL, "gevb_ctx_create: nccl_id required when nranks > 1");
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev == 0) GEVB_FAIL("gevb_ctx_create: no CUDA device available (%s); this library has no CPU fallback", cudaGetErrorString(e));
	GEVB_CHECK_ARG(device >= 0 && device < ndev, "gevb_ctx_create: device %d out of range (%d devices)", device, ndev);
	CUDA_TRY(cudaSetDevice(device));
	gevb_ctx * c = new gevb_ctx();
	memset(c, 0, sizeof(*c));
	c->N = ngrid; c->nh = ngrid / 2 + 1;
	c->device = device; c->rank = rank; c->nranks = nranks;
	GEVB_TRY(gevb_slab_geometry(ngrid, rank, nranks, &c->z0, &c->nzl, &c->ky0, &c->nkyl));
	cudaDeviceProp prop;
	CUDA_TRY(cudaGetDeviceProperties(&prop, device));
	c->num_sms = prop.multiProcessorCount;
	CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
	// k tables, computed on the host with the reference's own expressions (gevolution.hpp:222-227)
	std::vector<double> g(ngrid);
	std::vector<double2> ks(ngrid);
	for (int i = 0; i < ngrid; i++)
	{
		g[i] = 2. * (double) ngrid * sin(M_PI * (double) i / (double) ngrid);
		ks[i].x = g[i] * cos(M_PI * (double) i / (double) ngrid);
		ks[i].y = g[i] * -sin(M_PI * (double) i / (double) ngrid);
		g[i] *= g[i];
	}
	CUDA_TRY(cudaMalloc(&c->d_gridk2, sizeof(double) * ngrid));
	CUDA_TRY(cudaMalloc(&c->d_kshift, sizeof(double2) * ngrid));
	CUDA_TRY(cudaMemcpy(c->d_gridk2, g.data(), sizeof(double) * ngrid, cudaMemcpyHostToDevice));
	CUDA_TRY(cudaMemcpy(c->d_kshift, ks.data(), sizeof(double2) * ngrid, cudaMemcpyHostToDevice));
	CUDA_TRY(cudaMalloc(&c->d_red, sizeof(double) * 8192));
	CUDA_TRY(cudaMallocHost(&c->h_red, sizeof(double) * 8192));
	if (nranks > 1)
	{
		ncclUniqueId id;
		memcpy(&id, nccl_id, sizeof(id));
		GEVB_TRY(gevb_nccl_load());
		NCCL_TRY(ncclCommInitRank(&c->comm, nranks, id, rank));
		c->have_comm = true;
		GEVB_TRY(gevb_peer_setup(c));
	}
	*out = c;
	return 0;
}

extern "C" int gevb_ctx_destroy(gevb_ctx * c)
{
	if (c == NULL) return 0;
	cudaSetDevice(c->device);
	cudaStreamSynchronize(c->stream);
	gevb_xchg_release(c);
	gevb_peer_release(c);
	if (c->d_barrier)
	{
		cudaFree(c->d_barrier);
		cudaStreamDestroy(c->xstream);
		for (int k = 0; k <= GEVB_NXEV; k++) cudaEventDestroy(c->xev[k]);
	}
	if (c->have_comm) ncclCommDestroy(c->comm);
	cudaFree(c->d_gridk2); cudaFree(c->d_kshift); cudaFree(c->d_red); cudaFreeHost(c->h_red);
	if (c->scratch) cudaFree(c->scratch);
	if (c->scratch2) cudaFree(c->scratch2);
	cudaStreamDestroy(c->stream);
	if (c->timer) { for (size_t i = 0; i < c->timer->ev.size(); i++) cudaEventDestroy(c->timer->ev[i]); delete c->timer; }
	delete c;
	return 0;
}

extern "C" int gevb_ctx_sync(gevb_ctx * c)
{
	GEVB_CHECK_ARG(c != NULL, "gevb_ctx_sync: NULL context");
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	return 0;
}

extern "C" int gevb_ctx_geometry(gevb_ctx * c, int * ngrid, int * z0, int * nzl, int * ky0, int * nkyl)
{
	GEVB_CHECK_ARG(c != NULL, "gevb_ctx_geometry: NULL context");
	if (ngrid) *ngrid = c->N;
	if (z0) *z0 = c->z0;
	if (nzl) *nzl = c->nzl;
	if (ky0) *ky0 = c->ky0;
	if (nkyl) *nkyl = c->nkyl;
	return 0;
}

extern "C" void * gevb_ctx_stream(gevb_ctx * c) { return c ? (void *) c->stream : NULL; }
extern "C" int64_t gevb_ctx_launch_count(gevb_ctx * c) { return c ? c->launches : 0; }

static int grow(void ** buf, size_t * have, size_t want)
{
	if (*have >= want) return 0;
	if (*buf) CUDA_TRY(cudaFree(*buf));
	*buf = NULL; *have = 0;
	size_t sz = want + want / 8 + 256;
	CUDA_TRY(cudaMalloc(buf, sz));
	*have = sz;
	return 0;
}
int gevb_ctx_scratch(gevb_ctx * c, size_t bytes, void ** out)
{
	// the stream may still be using the old buffer
	if (c->scratch_bytes < bytes) CUDA_TRY(cudaStreamSynchronize(c->stream));
	GEVB_TRY(grow(&c->scratch, &c->scratch_bytes, bytes));
	*out = c->scratch;
	return 0;
}
int gevb_ctx_scratch2(gevb_ctx * c, size_t bytes, void ** out)
{
	if (c->scratch2_bytes < bytes) CUDA_TRY(cudaStreamSynchronize(c->stream));
	GEVB_TRY(grow(&c->scratch2, &c->scratch2_bytes, bytes));
	*out = c->scratch2;
	return 0;
}

// ---- parallel.sum / parallel.max --------------------------------------------
static int host_allreduce(gevb_ctx * c, double * v, int n, ncclRedOp_t op)
{
	GEVB_CHECK_ARG(c != NULL && v != NULL && n >= 0 && n <= 4096, "gevb_parallel_*: bad arguments");
	if (c->nranks == 1 || n == 0) return 0;
	CUDA_TRY(cudaSetDevice(c->device));
	memcpy(c->h_red, v, sizeof(double) * n);
	CUDA_TRY(cudaMemcpyAsync(c->d_red + 4096, c->h_red, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
	NCCL_TRY(ncclAllReduce(c->d_red + 4096, c->d_red + 4096, n, ncclDouble, op, c->comm, c->stream));
	CUDA_TRY(cudaMemcpyAsync(c->h_red, c->d_red + 4096, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	memcpy(v, c->h_red, sizeof(double) * n);
	return 0;
}
extern "C" int gevb_parallel_sum(gevb_ctx * c, double * v, int n) { return host_allreduce(c, v, n, ncclSum); }
extern "C" int gevb_parallel_max(gevb_ctx * c, double * v, int n) { return host_allreduce(c, v, n, ncclMax); }

// ---- fields ------------------------------------------------------------------
extern "C" int gevb_field_create(gevb_ctx * c, gevb_field ** out, int kind, int ncomp, int symmetric)
{
	GEVB_CHECK_ARG(c != NULL && out != NULL, "gevb_field_create: NULL argument");
	GEVB_CHECK_ARG(kind == GEVB_REAL || kind == GEVB_CPLX, "gevb_field_create: bad kind %d", kind);
	GEVB_CHECK_ARG(ncomp >= 1 && ncomp <= 16, "gevb_field_create: bad component count %d", ncomp);
	GEVB_CHECK_ARG(!symmetric || ncomp == 6, "gevb_field_create: symmetric fields are 3x3 (6 components)");
	GEVB_CHECK_ARG(kind == GEVB_REAL || (c->cplx_comp_stride() < (1ull << 31) && c->cplx_comp_stride() * (size_t) c->N < (1ull << 40)),
		"gevb_field_create: too many Fourier sites per rank for the 32-bit index decode (use more ranks)");
	CUDA_TRY(cudaSetDevice(c->device));
	gevb_field * f = new gevb_field();
	f->ctx = c; f->kind = kind; f->ncomp = ncomp; f->symmetric = symmetric;
	f->comp_stride = kind == GEVB_REAL ? c->real_comp_stride() : c->cplx_comp_stride();
	f->bytes = f->comp_stride * ncomp * (kind == GEVB_REAL ? sizeof(double) : sizeof(double2));
	cudaError_t e = cudaMalloc(&f->data, f->bytes);
	if (e != cudaSuccess) { delete f; GEVB_FAIL("gevb_field_create: cudaMalloc(%zu) failed: %s", f->bytes, cudaGetErrorString(e)); }
	CUDA_TRY(cudaMemsetAsync(f->data, 0, f->bytes, c->stream));
	*out = f;
	return 0;
}

extern "C" int gevb_field_destroy(gevb_field * f)
{
	if (f == NULL) return 0;
	cudaSetDevice(f->ctx->device);
	cudaStreamSynchronize(f->ctx->stream);
	cudaFree(f->data);
	delete f;
	return 0;
}

extern "C" int gevb_field_components(gevb_field * f) { return f ? f->ncomp : 0; }
extern "C" void * gevb_field_device_ptr(gevb_field * f) { return f ? (void *) f->data : NULL; }

extern "C" int gevb_field_upload(gevb_field * f, const double * host)
{
	GEVB_CHECK_ARG(f != NULL && host != NULL, "gevb_field_upload: NULL argument");
	gevb_ctx * c = f->ctx;
	CUDA_TRY(cudaSetDevice(c->device));
	if (f->kind == GEVB_REAL)
	{
		size_t bulk = (size_t) c->nzl * c->plane();
		for (int k = 0; k < f->ncomp; k++)
			CUDA_TRY(cudaMemcpyAsync(f->data + k * f->comp_stride + c->plane(), host + k * bulk, bulk * sizeof(double), cudaMemcpyHostToDevice, c->stream));
	}
	else
		CUDA_TRY(cudaMemcpyAsync(f->data, host, f->bytes, cudaMemcpyHostToDevice, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	return 0;
}

extern "C" int gevb_field_download(gevb_field * f, double * host)
{
	GEVB_CHECK_ARG(f != NULL && host != NULL, "gevb_field_download: NULL argument");
	gevb_ctx * c = f->ctx;
	CUDA_TRY(cudaSetDevice(c->device));
	if (f->kind == GEVB_REAL)
	{
		size_t bulk = (size_t) c->nzl * c->plane();
		for (int k = 0; k < f->ncomp; k++)
			CUDA_TRY(cudaMemcpyAsync(host + k * bulk, f->data + k * f->comp_stride + c->plane(), bulk * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	}
	else
		CUDA_TRY(cudaMemcpyAsync(host, f->data, f->bytes, cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	return 0;
}

// projection_init (main.cpp:378,426,438): zero everything, ghost planes included
extern "C" int gevb_projection_init(gevb_field * f)
{
	GEVB_CHECK_ARG(f != NULL, "projection_init: NULL field");
	CUDA_TRY(cudaSetDevice(f->ctx->device));
	Timed timed_(f->ctx, CLS_INIT);
	CUDA_TRY(cudaMemsetAsync(f->data, 0, f->bytes, f->ctx->stream));
	return 0;
}

// Field::updateHalo (main.cpp:518,568,598)
extern "C" int gevb_field_updateHalo(gevb_field * f)
{
	GEVB_CHECK_ARG(f != NULL && f->kind == GEVB_REAL, "updateHalo: needs a real field");
	gevb_ctx * c = f->ctx;
	CUDA_TRY(cudaSetDevice(c->device));
	Timed timed_(c, CLS_HALO);
	const size_t pl = c->plane();
	if (c->nranks == 1)
	{
		for (int k = 0; k < f->ncomp; k++)
		{
			double * base = f->data + k * f->comp_stride;
			CUDA_TRY(cudaMemcpyAsync(base, base + (size_t) c->nzl * pl, pl * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
			CUDA_TRY(cudaMemcpyAsync(base + (size_t) (c->nzl + 1) * pl, base + pl, pl * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
		}
		return 0;
	}
	if (gevb_peer_on(c) && (size_t) f->ncomp * pl * 2 <= c->pc_plane_doubles) return gevb_peer_halo(f);
	const int up = (c->rank + 1) % c->nranks, dn = (c->rank + c->nranks - 1) % c->nranks;
	NCCL_TRY(ncclGroupStart());
	for (int k = 0; k < f->ncomp; k++)
	{
		double * base = f->data + k * f->comp_stride;
		NCCL_TRY(ncclSend(base + pl, pl, ncclDouble, dn, c->comm, c->stream));                          // my first bulk plane -> lower neighbour's upper ghost
		NCCL_TRY(ncclSend(base + (size_t) c->nzl * pl, pl, ncclDouble, up, c->comm, c->stream));        // my last bulk plane  -> upper neighbour's lower ghost
		NCCL_TRY(ncclRecv(base + (size_t) (c->nzl + 1) * pl, pl, ncclDouble, up, c->comm, c->stream));
		NCCL_TRY(ncclRecv(base, pl, ncclDouble, dn, c->comm, c->stream));
	}
	NCCL_TRY(ncclGroupEnd());
	return 0;
}

__global__ void k_fold_add(double * __restrict__ dst, const double * __restrict__ src, size_t n, int ncomp, size_t dst_stride, size_t src_stride)
{
	for (int k = 0; k < ncomp; k++)
		for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
			dst[k * dst_stride + i] += src[k * src_stride + i];
}

// scalarProjectionCIC_comm / vectorProjectionCICNGP_comm / symtensorProjectionCICNGP_comm
// (gevolution.hpp:1024,1149,1300): deposits only reach x and +1 neighbours, x/y wrap is done
// by the deposit kernels, so the fold is one plane: upper ghost -> next rank's first bulk plane
extern "C" int gevb_projection_comm(gevb_field * f)
{
	GEVB_CHECK_ARG(f != NULL && f->kind == GEVB_REAL, "projection_comm: needs a real field");
	gevb_ctx * c = f->ctx;
	CUDA_TRY(cudaSetDevice(c->device));
	Timed timed_(c, CLS_COMM);
	const size_t pl = c->plane();
	const double * src = f->data + (size_t) (c->nzl + 1) * pl;
	size_t src_stride = f->comp_stride;
	if (c->nranks > 1 && gevb_peer_on(c) && (size_t) f->ncomp * pl <= c->pc_plane_doubles) return gevb_peer_fold(f);
	if (c->nranks > 1)
	{
		void * stage;
		GEVB_TRY(gevb_ctx_scratch(c, pl * f->ncomp * sizeof(double), &stage));
		const int up = (c->rank + 1) % c->nranks, dn = (c->rank + c->nranks - 1) % c->nranks;
		NCCL_TRY(ncclGroupStart());
		for (int k = 0; k < f->ncomp; k++)
		{
			NCCL_TRY(ncclSend(f->data + k * f->comp_stride + (size_t) (c->nzl + 1) * pl, pl, ncclDouble, up, c->comm, c->stream));
			NCCL_TRY(ncclRecv((double *) stage + k * pl, pl, ncclDouble, dn, c->comm, c->stream));
		}
		NCCL_TRY(ncclGroupEnd());
		src = (const double *) stage; src_stride = pl;
	}
	k_fold_add<<<gevb_grid(c, pl, 256), 256, 0, c->stream>>>(f->data + pl, src, pl, f->ncomp, f->comp_stride, src_stride);
	KERNEL_CHECK(c);
	return 0;
}

// ---- reductions ----------------------------------------------------------------
__global__ void k_block_sum(const double * __restrict__ v, size_t n, double * __restrict__ partial)
{
	__shared__ double sm[32];
	double s = 0.;
	for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) s += v[i];
	for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
	if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
	__syncthreads();
	if (threadIdx.x < 32)
	{
		s = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.;
		for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
		if (threadIdx.x == 0) partial[blockIdx.x] = s;
	}
}

// sum of the local bulk of one component, then parallel.sum (main.cpp:459-463)
extern "C" int gevb_field_sum(gevb_field * f, int comp, double * out)
{
	GEVB_CHECK_ARG(f != NULL && out != NULL && f->kind == GEVB_REAL, "gevb_field_sum: needs a real field");
	GEVB_CHECK_ARG(comp >= 0 && comp < f->ncomp, "gevb_field_sum: component %d out of range", comp);
	gevb_ctx * c = f->ctx;
	CUDA_TRY(cudaSetDevice(c->device));
	Timed timed_(c, CLS_SUM);
	const size_t n = (size_t) c->nzl * c->plane();
	int blocks = gevb_grid(c, n, 256, 4);
	if (blocks > 2048) blocks = 2048;
	k_block_sum<<<blocks, 256, 0, c->stream>>>(f->data + comp * f->comp_stride + c->plane(), n, c->d_red);
	KERNEL_CHECK(c);
	k_block_sum<<<1, 256, 0, c->stream>>>(c->d_red, blocks, c->d_red + 2048);
	KERNEL_CHECK(c);
	if (c->nranks > 1) NCCL_TRY(ncclAllReduce(c->d_red + 2048, c->d_red + 2048, 1, ncclDouble, ncclSum, c->comm, c->stream));
	CUDA_TRY(cudaMemcpyAsync(c->h_red, c->d_red + 2048, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	*out = c->h_red[0];
	return 0;
}

__global__ void k_add_constant(double * __restrict__ v, size_t n, double value)
{
	for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) v[i] += value;
}

extern "C" int gevb_field_add_constant(gevb_field * f, int comp, double value)
{
	GEVB_CHECK_ARG(f != NULL && f->kind == GEVB_REAL, "gevb_field_add_constant: needs a real field");
	GEVB_CHECK_ARG(comp >= 0 && comp < f->ncomp, "gevb_field_add_constant: component %d out of range", comp);
	gevb_ctx * c = f->ctx;
	CUDA_TRY(cudaSetDevice(c->device));
	const size_t n = (size_t) c->nzl * c->plane();
	k_add_constant<<<gevb_grid(c, n, 256), 256, 0, c->stream>>>(f->data + comp * f->comp_stride + c->plane(), n, value);
	KERNEL_CHECK(c);
	return 0;
}

__global__ void k_scale(double * __restrict__ v, size_t n, double factor)
{
	for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) v[i] *= factor;
}

// the site loops that rescale the stored vector potential around an output (output.hpp:212-218, hibernation.hpp:533-538,
// ic_read.hpp:312-317): every component of the local bulk times `factor`; the ghost planes are left to updateHalo
extern "C" int gevb_field_scale(gevb_field * f, double factor)
{
	GEVB_CHECK_ARG(f != NULL && f->kind == GEVB_REAL, "gevb_field_scale: needs a real field");
	gevb_ctx * c = f->ctx;
	CUDA_TRY(cudaSetDevice(c->device));
	const size_t n = (size_t) c->nzl * c->plane();
	for (int k = 0; k < f->ncomp; k++)
	{
		k_scale<<<gevb_grid(c, n, 256), 256, 0, c->stream>>>(f->data + k * f->comp_stride + c->plane(), n, factor);
		KERNEL_CHECK(c);
	}
	return 0;
}

extern "C" gevb_ctx * gevb_field_ctx(gevb_field * f) { return f ? f->ctx : NULL; }

// individual sites of one component (the IC generator's convolution kernel lives on 27 sites, ic_basic.hpp:737-1052)
extern "C" int gevb_field_set_sites(gevb_field * f, int comp, int n, const int * xyz, const double * values)
{
	GEVB_CHECK_ARG(f != NULL && f->kind == GEVB_REAL && n >= 0 && (n == 0 || (xyz != NULL && values != NULL)), "gevb_field_set_sites: bad arguments");
	GEVB_CHECK_ARG(comp >= 0 && comp < f->ncomp, "gevb_field_set_sites: component %d out of range", comp);
	gevb_ctx * c = f->ctx;
	CUDA_TRY(cudaSetDevice(c->device));
	for (int i = 0; i < n; i++)
	{
		const int x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
		GEVB_CHECK_ARG(x >= 0 && x < c->N && y >= 0 && y < c->N && z >= 0 && z < c->N, "gevb_field_set_sites: site (%d, %d, %d) outside the lattice", x, y, z);
		if (z < c->z0 || z >= c->z0 + c->nzl) continue;
		CUDA_TRY(cudaMemcpyAsync(f->data + comp * f->comp_stride + ((size_t) (z - c->z0 + 1) * c->N + y) * c->N + x, values + i, sizeof(double), cudaMemcpyHostToDevice, c->stream));
	}
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	return 0;
}
