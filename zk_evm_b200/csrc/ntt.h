// internal NTT interface (see ntt.cu)
#pragma once
#include "internal.h"
namespace zk {
// batched DIF: natural order in -> bit-reversed order out.  Transform t reads src + (t >> src_shift) * src_stride and
// writes dst + t * dst_stride (src may equal dst when src_shift == 0).  prescale tables multiply input element j.
void ntt_dif(Ctx& c, const uint64_t* src, size_t src_stride, unsigned src_shift, uint64_t* dst, size_t dst_stride,
             size_t ntrans, unsigned L, bool inverse, const uint64_t* prescale0, const uint64_t* prescale1,
             unsigned prescale_mask);
// out[col][bitrev(p)] = in[col][p] * scale * (tab ? tab[bitrev(p)] : 1)   (out of place)
void bitrev_permute(Ctx& c, const uint64_t* in, size_t in_stride, uint64_t* out, size_t out_stride, size_t ncols,
                    unsigned L, uint64_t scale, const uint64_t* tab);
// out[j] = c0 * base^j, j < len (uncached, caller-owned buffer)
void fill_powers(Ctx& c, uint64_t* out, size_t len, uint64_t base, uint64_t c0);
// cached table: out[j] = c0 * base^j, j < len
const uint64_t* get_power_table(Ctx& c, uint64_t base, uint64_t c0, size_t len);
// values (natural) -> coefficients (natural): ifft, or coset_ifft when coset_shift > 1
void intt_natural(Ctx& c, const uint64_t* values, uint64_t* scratch, uint64_t* coeffs, size_t ncols, unsigned L,
                  uint64_t coset_shift);
// coefficients (natural) -> evaluations on shift*<w_N> in bit-reversed order, N = n << rate_bits
void lde_bitrev(Ctx& c, const uint64_t* coeffs, uint64_t* lde, size_t ncols, unsigned L, unsigned rate_bits,
                uint64_t shift);
}  // namespace zk
