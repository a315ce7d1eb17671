// thermo.cuh -- K3 launch layer: batches of NucCruc problems (nuccruc.cuh) on the GPU.
//
// Data layout in HBM (struct of arrays, one problem per thread):
//   seq_a / seq_b   n x 32 bytes of base codes (A=0 C=1 G=2 T=3 I=4), 5'->3', zero (= A) padded: the padding is
//                   what the reference's ring buffer holds past the end of a sequence (see nuccruc.cuh header)
//   len_a / len_b   n bytes
//   log_strand      n floats: logf(total strand concentration), taken on the host with the reference's libm (NULL: every problem has
//                   the concentration whose logarithm is log_strand_all)
//   out             n x float4 {Tm, dH, dS, dG_dp}
//   order           optional permutation: thread `slot` works on problem order[slot] (problems binned by size, so that the lanes
//                   of a warp run DP fills of the same shape)
// The integer DP table (49 x 49 int32 = 9.6 KB, update_dp_param) is staged in shared memory once per CTA; the
// float parameter tables (58 KB, read sparsely by the evaluation epilogue) stay in global memory behind the
// read-only cache.  The DP matrix of a problem (one 16-bit word of trace bits and sign flags per cell, 2.2 KB) lives in the owning thread's
// local memory; all lanes of a warp walk the cells in the same (row, column) order, so those accesses coalesce.
#pragma once
#include "nuccruc.cuh"

#include <cuda_runtime.h>

namespace pcr {
namespace nc {

constexpr int THERMO_BLOCK = 128;
#ifndef THERMO_MIN_BLOCKS
#define THERMO_MIN_BLOCKS 5 // register budget of the strip-mined fill (nuccruc.cuh): 96 regs / thread
#endif
constexpr int THERMO_SEQ_STRIDE = 32;

struct DpShared {
	int dg[NPAIR * NPAIR];
};

__global__ void __launch_bounds__(THERMO_BLOCK, THERMO_MIN_BLOCKS) thermo_kernel(int op, uint32_t n, const uint32_t *__restrict__ order, const uint8_t *__restrict__ seq_a, const uint8_t *__restrict__ seq_b,
	const uint8_t *__restrict__ len_a, const uint8_t *__restrict__ len_b, const float *__restrict__ log_strand, const Tables *__restrict__ tables,
	const DpTable *__restrict__ dp, float4 *__restrict__ out, float log_strand_all = 0.0f)
{
	__shared__ DpTable s_dp;
	{
		const int *src = (const int *)dp;
		int *dst = (int *)&s_dp;
		for (int k = threadIdx.x; k < (int)(sizeof(DpTable) / sizeof(int)); k += blockDim.x) dst[k] = src[k];
	}
	__syncthreads();
	const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
	if (slot >= n) return;
	const uint32_t p = order ? order[slot] : slot; // size-binned launch order (thermo_abi.cu thermo_order)
	__align__(16) unsigned char q[NC_SEQ_CAP];
	__align__(16) unsigned char t[NC_SEQ_CAP];
	const uint4 *qa = (const uint4 *)(seq_a + (size_t)p * THERMO_SEQ_STRIDE);
	*(uint4 *)(q) = qa[0];
	*(uint4 *)(q + 16) = qa[1];
	*(uint32_t *)(q + 32) = 0u;
	const bool two = (op == OP_HETERODIMER || op == OP_HETERODIMER_DIAG);
	if (two) {
		const uint4 *tb = (const uint4 *)(seq_b + (size_t)p * THERMO_SEQ_STRIDE);
		*(uint4 *)(t) = tb[0];
		*(uint4 *)(t + 16) = tb[1];
		*(uint32_t *)(t + 32) = 0u;
	}
	__align__(16) unsigned short info[NC_CELLS];
	Ctx c;
	c.T = tables;
	c.D = &s_dp;
	c.q = q;
	c.t = two ? t : q;
	c.qlen = len_a[p];
	c.tlen = two ? len_b[p] : c.qlen;
	c.log_strand = log_strand ? log_strand[p] : log_strand_all;
	c.info = info + NC_INFO_PAD;
	const Result r = run_problem(c, op);
	out[p] = make_float4(r.tm, r.dH, r.dS, r.dp_dg);
}

// DP cells of one problem, as SURVEY.md section 8d counts them
__host__ __device__ inline long long problem_cells(int op, int qlen, int tlen)
{
	switch (op) {
	case OP_HAIRPIN: {
		const int s = qlen - 4;
		return s > 0 ? (long long)s * (s + 1) / 2 : 0;
	}
	case OP_HOMODIMER: return (long long)qlen * qlen;
	case OP_HETERODIMER: return (long long)qlen * tlen;
	case OP_HETERODIMER_DIAG: return qlen < tlen ? qlen : tlen;
	case OP_HOMODIMER_DIAG: return qlen;
	default: return 0;
	}
}

} // namespace nc
} // namespace pcr
