// snapshot.cu -- device side of the Gadget-2 snapshot writer (Particles_gevolution.hpp:30-251)
//
// One pass over a species: select the tracers (ID % tracer_factor == 0), apply the half-step corrections that
// EXACT_OUTPUT_REDSHIFTS adds (positions drifted by dtau_pos with the phi-corrected velocity, momenta kicked by
// dtau_vel with the CIC gradient of phi, Particles_gevolution.hpp:153-199), convert to Gadget units (kpc/h; km/s
// divided by sqrt(a)) in float32, and compact into contiguous output arrays.  The file itself is written by
// host/output.cpp.
#include "gevb_internal.cuh"

namespace {

#define GADGET_VELOCITY_CONVERSION 3.335640952e-6   // Gadget velocity unit / speed of light, metadata.hpp:117-118

__global__ void __launch_bounds__(256) k_gadget2(int64_t n, const double * __restrict__ x, const double * __restrict__ y, const double * __restrict__ z,
                                                  const double * __restrict__ qx, const double * __restrict__ qy, const double * __restrict__ qz, const int64_t * __restrict__ id,
                                                  const double * __restrict__ phi, int N, int z0, size_t plane, double a, double boxsize, int tracer_factor, double dtau_pos, double dtau_vel,
                                                  float * __restrict__ opos, float * __restrict__ ovel, int64_t * __restrict__ oid, unsigned long long * counter)
{
	const double dx = 1.0 / (double) N;
	const double rescale_vel = 1. / sqrt(a) / GADGET_VELOCITY_CONVERSION;             // :39
	for (int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x)
	{
		const int64_t pid = id[i];
		if (pid % tracer_factor != 0) continue;                                        // :152
		const double pos[3] = {x[i], y[i], z[i]}, vel[3] = {qx[i], qy[i], qz[i]};
		double phip = 0., gradphi[3] = {0., 0., 0.};
		if (phi != NULL)
		{
			double r[3], ip;
			for (int l = 0; l < 3; l++) r[l] = modf(pos[l] / dx, &ip);               // :157-158
			const int cx = cell_of(pos[0], dx, N), cy = cell_of(pos[1], dx, N), cz = cell_of(pos[2], dx, N) - z0;
			const int xp = cx + 1 == N ? 0 : cx + 1, yp = cy + 1 == N ? 0 : cy + 1;
			const double * p0 = phi + (size_t) (cz + 1) * plane, * p1 = p0 + plane;       // local plane cz and the one above (ghost plane at the slab top)
			const double f000 = p0[(size_t) cy * N + cx], f100 = p0[(size_t) cy * N + xp], f010 = p0[(size_t) yp * N + cx], f110 = p0[(size_t) yp * N + xp];
			const double f001 = p1[(size_t) cy * N + cx], f101 = p1[(size_t) cy * N + xp], f011 = p1[(size_t) yp * N + cx], f111 = p1[(size_t) yp * N + xp];
			phip = f000 * (1. - r[0]) * (1. - r[1]) * (1. - r[2]);                       // :160-167
			phip += f100 * r[0] * (1. - r[1]) * (1. - r[2]);
			phip += f010 * (1. - r[0]) * r[1] * (1. - r[2]);
			phip += f110 * r[0] * r[1] * (1. - r[2]);
			phip += f001 * (1. - r[0]) * (1. - r[1]) * r[2];
			phip += f101 * r[0] * (1. - r[1]) * r[2];
			phip += f011 * (1. - r[0]) * r[1] * r[2];
			phip += f111 * r[0] * r[1] * r[2];
			gradphi[0] = (1. - r[1]) * (1. - r[2]) * (f100 - f000);                      // :169-180
			gradphi[1] = (1. - r[0]) * (1. - r[2]) * (f010 - f000);
			gradphi[2] = (1. - r[0]) * (1. - r[1]) * (f001 - f000);
			gradphi[0] += r[1] * (1. - r[2]) * (f110 - f010);
			gradphi[1] += r[0] * (1. - r[2]) * (f110 - f100);
			gradphi[2] += r[0] * (1. - r[1]) * (f101 - f100);
			gradphi[0] += (1. - r[1]) * r[2] * (f101 - f001);
			gradphi[1] += (1. - r[0]) * r[2] * (f011 - f001);
			gradphi[2] += (1. - r[0]) * r[1] * (f011 - f010);
			gradphi[0] += r[1] * r[2] * (f111 - f011);
			gradphi[1] += r[0] * r[2] * (f111 - f101);
			gradphi[2] += r[0] * r[1] * (f111 - f110);
		}
		double w0 = vel[0] * vel[0] + vel[1] * vel[1] + vel[2] * vel[2];              // :183-187 (ref_dist[] reused as scratch there)
		double w1 = w0 + a * a;
		const double w2 = sqrt(w1);
		w0 += w1;
		w1 = 1. + (4. - (w0 / w1)) * phip;
		const unsigned long long slot = atomicAdd(counter, 1ull);
		double ip;
		for (int l = 0; l < 3; l++)
		{
			opos[3 * slot + l] = (float) (modf(1. + pos[l] + dtau_pos * vel[l] * w1 / w2, &ip) * boxsize);           // :189-190
			ovel[3 * slot + l] = (float) ((vel[l] - dtau_vel * w0 * gradphi[l] / dx / w2) * rescale_vel / a);       // :192-193
		}
		oid[slot] = pid;
	}
}

} // namespace

extern "C" gevb_ctx * gevb_pcls_ctx(gevb_pcls * p) { return p ? p->ctx : NULL; }

extern "C" int gevb_ctx_ranks(gevb_ctx * c, int * rank, int * nranks)
{
	GEVB_CHECK_ARG(c != NULL, "gevb_ctx_ranks: NULL context");
	if (rank) *rank = c->rank;
	if (nranks) *nranks = c->nranks;
	return 0;
}

extern "C" int gevb_pcls_gadget2_arrays(gevb_pcls * p, double a, double boxsize, int tracer_factor, double dtau_pos, double dtau_vel, gevb_field * phi,
                                         float * pos, float * vel, int64_t * ids, int64_t * n_out)
{
	GEVB_CHECK_ARG(p != NULL && n_out != NULL, "gevb_pcls_gadget2_arrays: NULL argument");
	GEVB_CHECK_ARG(tracer_factor >= 1, "gevb_pcls_gadget2_arrays: tracer factor must be >= 1");
	GEVB_CHECK_ARG(phi == NULL || (phi->kind == GEVB_REAL && phi->ncomp == 1 && phi->ctx == p->ctx), "gevb_pcls_gadget2_arrays: phi must be a one-component real field of the same context");
	gevb_ctx * c = p->ctx;
	*n_out = 0;
	if (p->n == 0) return 0;
	GEVB_CHECK_ARG(pos != NULL && vel != NULL && ids != NULL, "gevb_pcls_gadget2_arrays: NULL output array");
	CUDA_TRY(cudaSetDevice(c->device));
	void * stage;
	GEVB_TRY(gevb_ctx_scratch(c, (size_t) p->n * 32, &stage));
	int64_t * dids = (int64_t *) stage;
	float * dpos = (float *) (dids + p->n), * dvel = dpos + 3 * p->n;
	unsigned long long * counter = (unsigned long long *) (c->d_red + 4000);
	CUDA_TRY(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), c->stream));
	const int b = p->cur;
	k_gadget2<<<gevb_grid(c, (size_t) p->n, 256), 256, 0, c->stream>>>(p->n, p->x[b], p->y[b], p->z[b], p->qx[b], p->qy[b], p->qz[b], p->id[b],
		phi ? phi->data : NULL, c->N, c->z0, c->plane(), a, boxsize, tracer_factor, dtau_pos, dtau_vel, dpos, dvel, dids, counter);
	KERNEL_CHECK(c);
	unsigned long long nsel = 0;
	CUDA_TRY(cudaMemcpyAsync(&nsel, counter, sizeof(nsel), cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	CUDA_TRY(cudaMemcpyAsync(ids, dids, sizeof(int64_t) * nsel, cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaMemcpyAsync(pos, dpos, sizeof(float) * 3 * nsel, cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaMemcpyAsync(vel, dvel, sizeof(float) * 3 * nsel, cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	*n_out = (int64_t) nsel;
	return 0;
}
