// tests/native/host_sw_harness.cpp -- TEST INFRASTRUCTURE: the product's Smith-Waterman core
// (pcramp_b200/csrc/sw.cuh, __host__ __device__) compiled for the host so the CPU tier can check it against the
// reference's SO::SeqOverlap goldens without a GPU.  Never linked into libpcramp_gpu.so.
#include <algorithm>
using std::max;
using std::min;
#include "../../pcramp_b200/csrc/sw.cuh"

using namespace pcr;

extern "C" {
// out: n x 6 int32 {score, q_start, q_stop, t_start, t_stop, last_two}; coordinates are -1 when no cell reached 0
int host_sw_batch(unsigned n, const uint64_t *query, const uint64_t *target, int with_start, int *out)
{
	for (unsigned p = 0; p < n; ++p) {
		W128 qw, tw;
		qw.hi = query[2 * p]; qw.lo = query[2 * p + 1];
		tw.hi = target[2 * p]; tw.lo = target[2 * p + 1];
		sw::Query q;
		sw::query_from_word(qw, q);
		const sw::WordTarget t(tw);
		// every second problem through the instantiation a warp would pick for it, the others through the full-height one
		const int rows = (p & 1u) ? q.len : sw::SW_MAX_QUERY;
		const sw::Result r = with_start ? sw::align_rows<true>(q, t, rows) : sw::align_rows<false>(q, t, rows);
		unsigned a, b;
		sw::last_two(r, t, a, b);
		int *o = out + 6 * (size_t)p;
		o[0] = r.score;
		o[1] = r.any && with_start ? r.q_start : -1;
		o[2] = r.any ? r.q_stop : -1;
		o[3] = r.any && with_start ? r.t_start : -1;
		o[4] = r.any ? r.t_stop : -1;
		o[5] = r.any ? (int)((a << 4) | b) : -1;
	}
	return 0;
}
}
