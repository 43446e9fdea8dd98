// source_kernels.cu -- real-space source preparation, one HBM pass each
//
//   prepareFTsource (scalar)  gevolution.hpp:170-192   32 B / site (phi, chi, source in; result out)
//   prepareFTsource (tensor)  gevolution.hpp:57-147   104 B / site (phi, 6 Tij in; 6 Sij out)
//
// PHINONLINEAR on, ORIGINALMETRIC off (reference makefile:21).  One thread per
// site, x fastest so every load is a coalesced row segment; the phi stencil
// neighbours (x+-1 in the row, y+-1 rows, z+-1 planes incl. ghost planes) are
// served from L1/L2, so DRAM sees each phi value once.  x and y wrap by index
// arithmetic; z uses the ghost planes filled by updateHalo.
#include "gevb_internal.cuh"

namespace {

struct RGeom { int N, nzl; size_t plane; };

__global__ void __launch_bounds__(256) k_prepare_scalar(RGeom G, const double * __restrict__ phi, const double * __restrict__ chi, const double * source, double bgmodel, double * result, double coeff, double coeff2, double coeff3)
{
	const int N = G.N;
	const size_t total = (size_t) G.nzl * G.plane;
	for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < total; i += (size_t) gridDim.x * blockDim.x)
	{
		const int x = (int) (i % N); const size_t r = i / N;
		const int y = (int) (r % N); const int zl = (int) (r / N);
		const size_t row = ((size_t) (zl + 1) * N + y) * N;
		const int xm = x == 0 ? N - 1 : x - 1, xp = x == N - 1 ? 0 : x + 1;
		const size_t rowm = ((size_t) (zl + 1) * N + (y == 0 ? N - 1 : y - 1)) * N;
		const size_t rowp = ((size_t) (zl + 1) * N + (y == N - 1 ? 0 : y + 1)) * N;
		const double p = phi[row + x];
		double res = coeff2 * (__ldcs(source + row + x) - bgmodel);                     // :176
		res *= 1. - 2. * p;                                                             // :184
		const double d0 = phi[row + xm] - phi[row + xp];
		const double d1 = phi[rowm + x] - phi[rowp + x];
		const double d2 = phi[row + x - G.plane] - phi[row + x + G.plane];
		res += 0.125 * d0 * d0;                                                         // :185
		res += 0.125 * d1 * d1;                                                         // :186
		res += 0.125 * d2 * d2;                                                         // :187
		res += (coeff3 - coeff) * p - coeff3 * __ldcs(chi + row + x);                   // :190
		__stcs(result + row + x, res);
	}
}

__global__ void __launch_bounds__(256) k_prepare_tensor(RGeom G, const double * __restrict__ phi, const double * T, double * S, size_t cs, double coeff)
{
	const int N = G.N;
	const size_t total = (size_t) G.nzl * G.plane;
	for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < total; i += (size_t) gridDim.x * blockDim.x)
	{
		const int x = (int) (i % N); const size_t r = i / N;
		const int y = (int) (r % N); const int zl = (int) (r / N);
		const int xm = x == 0 ? N - 1 : x - 1, xp = x == N - 1 ? 0 : x + 1;
		const int ym = y == 0 ? N - 1 : y - 1, yp = y == N - 1 ? 0 : y + 1;
		const size_t pl = G.plane;
		const size_t row = ((size_t) (zl + 1) * N + y) * N, rowm = ((size_t) (zl + 1) * N + ym) * N, rowp = ((size_t) (zl + 1) * N + yp) * N;
		const size_t s = row + x;
		const double p0 = phi[s];
		const double px = phi[row + xp], mx = phi[row + xm];
		const double py = phi[rowp + x], my = phi[rowm + x];
		const double pz = phi[s + pl], mz = phi[s - pl];
		const double pxy = phi[rowp + xp], pxz = phi[row + xp + pl], pyz = phi[rowp + x + pl];
		double v;
		v = coeff * __ldcs(T + 0 * cs + s); v += 0.5 * (px - mx) * (px - mx); __stcs(S + 0 * cs + s, v);          // :64,70
		v = coeff * __ldcs(T + 3 * cs + s); v += 0.5 * (py - my) * (py - my); __stcs(S + 3 * cs + s, v);          // :75,81
		v = coeff * __ldcs(T + 5 * cs + s); v += 0.5 * (pz - mz) * (pz - mz); __stcs(S + 5 * cs + s, v);          // :86,92
		v = coeff * __ldcs(T + 1 * cs + s);                                                                     // :97
		v += px * py - p0 * pxy;                                                                               // :99
		v += 0.5 * p0 * p0; v -= 0.5 * px * px; v -= 0.5 * py * py; v += 0.5 * pxy * pxy;                      // :106-109
		__stcs(S + 1 * cs + s, v);
		v = coeff * __ldcs(T + 2 * cs + s);                                                                     // :114
		v += px * pz - p0 * pxz;                                                                               // :116
		v += 0.5 * p0 * p0; v -= 0.5 * px * px; v -= 0.5 * pz * pz; v += 0.5 * pxz * pxz;                      // :123-126
		__stcs(S + 2 * cs + s, v);
		v = coeff * __ldcs(T + 4 * cs + s);                                                                     // :131
		v += py * pz - p0 * pyz;                                                                               // :133
		v += 0.5 * p0 * p0; v -= 0.5 * py * py; v -= 0.5 * pz * pz; v += 0.5 * pyz * pyz;                      // :140-143
		__stcs(S + 4 * cs + s, v);
	}
}

int check_real(const gevb_field * f, int ncomp, const char * who, const char * name)
{
	GEVB_CHECK_ARG(f != NULL, "%s: %s is NULL", who, name);
	GEVB_CHECK_ARG(f->kind == GEVB_REAL, "%s: %s must be a real-space field", who, name);
	GEVB_CHECK_ARG(f->ncomp == ncomp, "%s: %s needs %d components (has %d)", who, name, ncomp, f->ncomp);
	return 0;
}

} // namespace

extern "C" int gevb_prepareFTsource_scalar(gevb_field * phi, gevb_field * chi, gevb_field * source, double bgmodel, gevb_field * result, double coeff, double coeff2, double coeff3)
{
	GEVB_TRY(check_real(phi, 1, "prepareFTsource", "phi"));
	GEVB_TRY(check_real(chi, 1, "prepareFTsource", "chi"));
	GEVB_TRY(check_real(source, 1, "prepareFTsource", "source"));
	GEVB_TRY(check_real(result, 1, "prepareFTsource", "result"));
	GEVB_CHECK_ARG(result != phi && result != chi, "prepareFTsource: result must not alias phi or chi");
	gevb_ctx * c = phi->ctx;
	CUDA_TRY(cudaSetDevice(c->device));
	Timed timed_(c, CLS_PREP_SCALAR);
	RGeom G = {c->N, c->nzl, c->plane()};
	k_prepare_scalar<<<gevb_grid(c, (size_t) c->nzl * c->plane(), 256), 256, 0, c->stream>>>(G, phi->data, chi->data, source->data, bgmodel, result->data, coeff, coeff2, coeff3);
	KERNEL_CHECK(c);
	return 0;
}

extern "C" int gevb_prepareFTsource_tensor(gevb_field * phi, gevb_field * Tij, gevb_field * Sij, double coeff)
{
	GEVB_TRY(check_real(phi, 1, "prepareFTsource", "phi"));
	GEVB_TRY(check_real(Tij, 6, "prepareFTsource", "Tij"));
	GEVB_TRY(check_real(Sij, 6, "prepareFTsource", "Sij"));
	gevb_ctx * c = phi->ctx;
	CUDA_TRY(cudaSetDevice(c->device));
	Timed timed_(c, CLS_PREP_TENSOR);
	RGeom G = {c->N, c->nzl, c->plane()};
	k_prepare_tensor<<<gevb_grid(c, (size_t) c->nzl * c->plane(), 256), 256, 0, c->stream>>>(G, phi->data, Tij->data, Sij->data, Sij->comp_stride, coeff);
	KERNEL_CHECK(c);
	return 0;
}
