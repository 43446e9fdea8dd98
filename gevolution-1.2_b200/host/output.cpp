// output.cpp -- the outputs the parity metrics are defined on
//
//   writePowerSpectrum        tools.hpp:268-346     text file, EXACT_OUTPUT_REDSHIFTS interpolation between two calls
//   writeSpectra (phi, chi, hij, B)   output.hpp:1945-1981, 2151-2155   the call sequence around extractPowerSpectrum
//   saveGadget2               Particles_gevolution.hpp:30-251   Gadget-2 binary snapshot (float32 positions in kpc/h,
//                             velocities in km/s / sqrt(a), int64 IDs), with the half-step position / velocity
//                             corrections of EXACT_OUTPUT_REDSHIFTS evaluated on the device
//
// The device work (FFT, projection, binning, unit conversion of the particles) goes through the C ABI; this file is
// the host side: file formats and the order of calls.  Written against include/gevolution_b200.hpp.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#define GEVB_THROW_ON_ERROR
#include "../../include/gevolution_b200.hpp"
#include "background.hpp"

using namespace gevb200;

// ---------------------------------------------------------------------------------------------------------------
// tools.hpp:268-346.  File layout (one line per occupied bin):
//   # <description>
//   # redshift z=<z>
//   # k              Pk             sigma(k)       sigma(Pk)      count
//     k/rescalek   P/rescalep   kscatter/rescalek   pscatter/rescalep/sqrt(count)   count
// If the file already exists and the scale factor has passed the target redshift, the spectrum written is the linear
// interpolation in redshift between the stored one (taken just before the target) and the current one.
extern "C" int gevb_writePowerSpectrum(const double * kbin, const double * power, const double * kscatter, const double * pscatter, const int * occupation, int numbins,
                                        double rescalek, double rescalep, const char * filename, const char * description, double a, double z_target)
{
	if (!kbin || !power || !kscatter || !pscatter || !occupation || !filename || !description || numbins < 1) return 1;
	std::vector<double> out(numbins);
	for (int i = 0; i < numbins; i++) out[i] = power[i] / rescalep;
	if (1. / a < z_target + 1.)
	{
		if (FILE * prev = std::fopen(filename, "r"))
		{
			char line[512];
			double z_prev = 0.;
			bool ok = std::fgets(line, sizeof(line), prev) != NULL;                       // description
			ok = ok && std::fgets(line, sizeof(line), prev) != NULL && std::sscanf(line, "# redshift z=%lf", &z_prev) == 1;
			ok = ok && std::fgets(line, sizeof(line), prev) != NULL;                       // column names
			if (!ok) std::fprintf(stderr, " error parsing power spectrum file header for interpolation (EXACT_OUTPUT_REDSHIFTS)\n");
			else
			{
				std::vector<double> stored(out);
				for (int i = 0; i < numbins && ok; i++)
				{
					if (occupation[i] <= 0) continue;
					double k_, p_;
					ok = std::fgets(line, sizeof(line), prev) != NULL && std::sscanf(line, " %le %le", &k_, &p_) == 2;
					if (ok) stored[i] = p_;
					else std::fprintf(stderr, " error parsing power spectrum file data %d for interpolation (EXACT_OUTPUT_REDSHIFTS)\n", i);
				}
				const double weight = (z_prev - z_target) / (1. + z_prev - 1. / a);        // tools.hpp:293
				for (int i = 0; i < numbins; i++) out[i] = (1. - weight) * stored[i] + weight * power[i] / rescalep;
				a = 1. / (z_target + 1.);
			}
			std::fclose(prev);
		}
	}
	FILE * f = std::fopen(filename, "w");
	if (f == NULL) { std::fprintf(stderr, " error opening file for power spectrum output!\n"); return 1; }
	std::fprintf(f, "# %s\n", description);
	std::fprintf(f, "# redshift z=%f\n", (1. / a) - 1.);
	std::fprintf(f, "# k              Pk             sigma(k)       sigma(Pk)      count\n");
	for (int i = 0; i < numbins; i++)
		if (occupation[i] > 0)
			std::fprintf(f, "  %e   %e   %e   %e   %d\n", kbin[i] / rescalek, out[i], kscatter[i] / rescalek, pscatter[i] / rescalep / std::sqrt((double) occupation[i]), occupation[i]);
	std::fclose(f);
	return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Gadget-2 snapshot of one species (Particles_gevolution.hpp:30-251, GADGET_ID_BYTES == 8).
// header256: the caller's 256-byte gadget2_header (metadata.hpp:152-171); npart[1] is overwritten with the number of
// particles written (ID % tracer_factor == 0), time = a and BoxSize (kpc/h) are read from it as the reference does.
// All ranks write their share into the one file (num_files == 1) at the offsets the reference computes.
extern "C" int gevb_pcls_saveGadget2(gevb_pcls * p, const char * filename, void * header256, int tracer_factor, double dtau_pos, double dtau_vel, gevb_field * phi)
{
	if (p == NULL || filename == NULL || header256 == NULL || tracer_factor < 1) return 1;
	unsigned char * hdr = (unsigned char *) header256;
	double time, boxsize;
	std::memcpy(&time, hdr + 6 * 4 + 6 * 8, 8);                                          // gadget2_header::time
	std::memcpy(&boxsize, hdr + 6 * 4 + 6 * 8 + 2 * 8 + 2 * 4 + 6 * 4 + 2 * 4, 8);        // gadget2_header::BoxSize
	int64_t n_local = 0;
	if (gevb_pcls_count(p, &n_local) != 0) return 1;
	std::vector<float> pos(3 * (size_t) n_local), vel(3 * (size_t) n_local);
	std::vector<int64_t> ids((size_t) n_local);
	int64_t n_sel = 0;
	if (gevb_pcls_gadget2_arrays(p, time, boxsize, tracer_factor, dtau_pos, dtau_vel, phi, pos.data(), vel.data(), ids.data(), &n_sel) != 0) return 1;
	// particles written by the lower ranks, and in total
	int rank = 0, nranks = 1;
	gevb_ctx * ctx = gevb_pcls_ctx(p);
	gevb_ctx_ranks(ctx, &rank, &nranks);
	std::vector<double> counts(nranks, 0.);
	counts[rank] = (double) n_sel;
	if (gevb_parallel_sum(ctx, counts.data(), nranks) != 0) return 1;
	uint64_t before = 0, total = 0;
	for (int r = 0; r < nranks; r++) { if (r < rank) before += (uint64_t) counts[r]; total += (uint64_t) counts[r]; }
	const uint32_t npart1 = (uint32_t) total;
	std::memcpy(hdr + 4, &npart1, 4);                                                     // hdr.npart[1]
	// [4][hdr 256][4] [4][pos 12 n][4] [4][vel 12 n][4] [4][ids 8 n][4]
	const uint64_t off_pos = 4 + 256 + 4 + 4, off_vel = off_pos + 12 * total + 4 + 4, off_id = off_vel + 12 * total + 4 + 4;
	bool ok = true;
	FILE * f = NULL;
	auto put = [&](uint64_t off, const void * data, size_t bytes) { ok = ok && fseeko(f, (off_t) off, SEEK_SET) == 0 && (bytes == 0 || std::fwrite(data, 1, bytes, f) == bytes); };
	if (rank == 0)
	{
		// rank 0 creates the file and writes the header and the Fortran-style block markers
		f = std::fopen(filename, "wb");
		if (f == NULL) { std::fprintf(stderr, " error opening %s for Gadget-2 output\n", filename); ok = false; }
		else
		{
			uint32_t b = 256;
			put(0, &b, 4); put(4, hdr, 256); put(260, &b, 4);
			b = (uint32_t) (12 * total); put(off_pos - 4, &b, 4); put(off_vel - 8, &b, 4); put(off_vel - 4, &b, 4); put(off_id - 8, &b, 4);
			b = (uint32_t) (8 * total); put(off_id - 4, &b, 4); put(off_id + 8 * total, &b, 4);
			std::fclose(f); f = NULL;
		}
	}
	double created = ok ? 0. : 1.;
	if (gevb_parallel_sum(ctx, &created, 1) != 0 || created != 0.) return 1;             // the file exists on every rank's view from here on
	f = std::fopen(filename, "r+b");
	if (f == NULL) { std::fprintf(stderr, " error opening %s for Gadget-2 output\n", filename); ok = false; }
	else
	{
		put(off_pos + 12 * before, pos.data(), 12 * (size_t) n_sel);
		put(off_vel + 12 * before, vel.data(), 12 * (size_t) n_sel);
		put(off_id + 8 * before, ids.data(), 8 * (size_t) n_sel);
		std::fclose(f);
	}
	// every rank has finished writing when the next collective completes
	double done = ok ? 0. : 1.;
	if (gevb_parallel_sum(ctx, &done, 1) != 0) return 1;
	return done == 0. ? 0 : 1;
}
