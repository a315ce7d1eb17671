// tests/native/host_thermo_harness.cpp -- TEST INFRASTRUCTURE.
//
// Compiles the product's NucCruc device code (pcramp_b200/csrc/nuccruc.cuh, every function is
// __host__ __device__) for the HOST so that the CPU-only test tier can check it against the reference's
// golden vectors and, in the dev container, against the live compiled reference -- there is no GPU in the
// dev container.  Never linked into libpcramp_gpu.so; the product only runs the CUDA kernels.
#include "../../pcramp_b200/csrc/nuccruc.cuh"
#include <stdint.h>
#include <vector>

using namespace pcr::nc;

static Tables g_tables;
static bool g_init = false;

extern "C" {

// seqs: n problems, 2 strings each of stride 33 bytes (NUL padded); out: n x 4 floats {tm, dH, dS, dp_dg}
int host_thermo_batch(int op, int n, const char *a, const char *b, float salt, const float *strand, float *out, long long *cells)
{
	if (!g_init) {
		build_tables(g_tables);
		g_init = true;
	}
	DpTable dp;
	build_dp(g_tables, salt, 310.15f, dp);
	std::vector<unsigned short> info(NC_CELLS);
	long long total = 0;
	for (int p = 0; p < n; ++p) {
		unsigned char q[NC_SEQ_CAP], t[NC_SEQ_CAP];
		memset(q, 0, sizeof(q));
		memset(t, 0, sizeof(t));
		const char *sa = a + (size_t)p * 33, *sb = b ? b + (size_t)p * 33 : sa;
		int qlen = 0, tlen = 0;
		for (; qlen < 32 && sa[qlen]; ++qlen) {
			const int c = base_code(sa[qlen]);
			if (c < 0) return -1;
			q[qlen] = (unsigned char)c;
		}
		for (; tlen < 32 && sb[tlen]; ++tlen) {
			const int c = base_code(sb[tlen]);
			if (c < 0) return -1;
			t[tlen] = (unsigned char)c;
		}
		Ctx c;
		c.T = &g_tables;
		c.D = &dp;
		c.q = q;
		c.t = (op == OP_HETERODIMER || op == OP_HETERODIMER_DIAG) ? t : q;
		c.qlen = qlen;
		c.tlen = (op == OP_HETERODIMER || op == OP_HETERODIMER_DIAG) ? tlen : qlen;
		c.log_strand = logf(strand[p]);
		c.info = info.data() + NC_INFO_PAD;
		Result r = run_problem(c, op);
		out[4 * p + 0] = r.tm;
		out[4 * p + 1] = r.dH;
		out[4 * p + 2] = r.dS;
		out[4 * p + 3] = r.dp_dg;
		total += r.cells;
	}
	if (cells) *cells = total;
	return 0;
}

void host_dp_table(float salt, int *out)
{
	if (!g_init) {
		build_tables(g_tables);
		g_init = true;
	}
	DpTable dp;
	build_dp(g_tables, salt, 310.15f, dp);
	memcpy(out, dp.dg, sizeof(dp.dg));
}
}
