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
	// one file with 32-bit block markers (MASK_MULTI is not built): the reference refuses such a write too (output.hpp:404)
	if (12 * total >= (1ull << 32)) { std::fprintf(stderr, " error: %llu particles do not fit one Gadget-2 file (32-bit block markers)\n", (unsigned long long) total); return 1; }
	const uint32_t npart1 = (uint32_t) total;
	uint32_t requested;
	std::memcpy(&requested, hdr + 4, 4);
	// the reference reports the mismatch and writes what it found (Particles_gevolution.hpp:90,108); npartTotal stays the caller's
	if (rank == 0 && requested != npart1) std::fprintf(stderr, " error: number of particles in saveGadget2 does not match request!\n");
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

// ---------------------------------------------------------------------------------------------------------------
// Field::saveHDF5 / loadHDF5 stand-ins (output.hpp:98-300; hibernation.hpp:588-600; ic_read.hpp:310,326).  HDF5 is not
// available in this image, so the dataset is a flat binary file: a 32-byte header, then the whole lattice as
// float64 [comp][z][y][x].  Every rank writes / reads the planes of its own slab at their global offsets.
namespace {
struct RawHeader { char magic[8]; int32_t ngrid, ncomp; char pad[16]; };
const char RAW_MAGIC[8] = {'G', 'E', 'V', 'B', 'F', 'L', 'D', '1'};
}

extern "C" int gevb_field_save_raw(gevb_field * f, const char * filename)
{
	if (f == NULL || filename == NULL) return 1;
	gevb_ctx * ctx = gevb_field_ctx(f);
	int rank = 0, nranks = 1, n, z0, nzl;
	gevb_ctx_ranks(ctx, &rank, &nranks);
	gevb_ctx_geometry(ctx, &n, &z0, &nzl, NULL, NULL);
	const int ncomp = gevb_field_components(f);
	const size_t slab = (size_t) nzl * n * n, comp_sites = (size_t) n * n * n;
	bool ok = true;
	std::vector<double> buf;
	try { buf.resize(slab * ncomp); } catch (...) { ok = false; }
	ok = ok && gevb_field_download(f, buf.data()) == 0;
	if (rank == 0 && ok)
	{
		FILE * h = std::fopen(filename, "wb");
		RawHeader hd;
		std::memset(&hd, 0, sizeof(hd));
		std::memcpy(hd.magic, RAW_MAGIC, 8); hd.ngrid = n; hd.ncomp = ncomp;
		ok = h != NULL && std::fwrite(&hd, 1, sizeof(hd), h) == sizeof(hd);
		if (h) ok = std::fclose(h) == 0 && ok;
		if (!ok) std::fprintf(stderr, " error opening %s for field output\n", filename);
	}
	double created = ok ? 0. : 1.;
	if (gevb_parallel_sum(ctx, &created, 1) != 0 || created != 0.) return 1;             // the file exists for every rank from here on
	FILE * h = std::fopen(filename, "r+b");
	ok = h != NULL;
	for (int k = 0; k < ncomp && ok; k++)
		ok = fseeko(h, (off_t) (sizeof(RawHeader) + ((size_t) k * comp_sites + (size_t) z0 * n * n) * sizeof(double)), SEEK_SET) == 0
			&& std::fwrite(buf.data() + (size_t) k * slab, sizeof(double), slab, h) == slab;
	if (h) ok = std::fclose(h) == 0 && ok;
	double done = ok ? 0. : 1.;
	if (gevb_parallel_sum(ctx, &done, 1) != 0) return 1;
	return done == 0. ? 0 : 1;
}

extern "C" int gevb_field_load_raw(gevb_field * f, const char * filename)
{
	if (f == NULL || filename == NULL) return 1;
	gevb_ctx * ctx = gevb_field_ctx(f);
	int n, z0, nzl;
	gevb_ctx_geometry(ctx, &n, &z0, &nzl, NULL, NULL);
	const int ncomp = gevb_field_components(f);
	const size_t slab = (size_t) nzl * n * n, comp_sites = (size_t) n * n * n;
	// read and validate everything on the host first; the ranks agree before anyone touches the device (no rank may
	// skip a collective the others enter)
	bool ok = true;
	std::vector<double> buf;
	try { buf.resize(slab * ncomp); } catch (...) { ok = false; }
	FILE * h = ok ? std::fopen(filename, "rb") : NULL;
	RawHeader hd;
	ok = ok && h != NULL && std::fread(&hd, 1, sizeof(hd), h) == sizeof(hd) && std::memcmp(hd.magic, RAW_MAGIC, 8) == 0 && hd.ngrid == n && hd.ncomp == ncomp;
	for (int k = 0; k < ncomp && ok; k++)
		ok = fseeko(h, (off_t) (sizeof(RawHeader) + ((size_t) k * comp_sites + (size_t) z0 * n * n) * sizeof(double)), SEEK_SET) == 0
			&& std::fread(buf.data() + (size_t) k * slab, sizeof(double), slab, h) == slab;
	if (h) std::fclose(h);
	double bad = ok ? 0. : 1.;
	if (gevb_parallel_sum(ctx, &bad, 1) != 0 || bad != 0.) return 1;
	if (gevb_field_upload(f, buf.data()) != 0) return 1;
	return gevb_field_updateHalo(f);
}
