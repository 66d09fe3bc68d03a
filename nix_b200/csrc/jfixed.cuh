// jfixed.cuh -- arithmetic of the NIX_J_FIXED experiment of push_deposit.cu (k_deposit's J tile as 64-bit fixed-point
// words added to with two native 32-bit shared-memory atomics).  EXPERIMENT: compiled only with -DNIX_J_FIXED=1, never
// run on a device (prepared after the round's GPU budget was spent).  The three functions below are
// plain arithmetic around atomicAdd(unsigned*, unsigned); tests/test_jfixed_arith.py compiles THIS file for the host with
// one-line stand-ins for the CUDA intrinsics and checks exactness, order independence, carries and signs.
#pragma once

// 2^(44 - exponent of the largest |q|, |q del/dt|): every per-particle contribution (|weight| <= 1) lands below 2^45,
// which leaves 2^18 of them room in a 64-bit word; scaling by a power of two is exact
template <typename T>
__device__ __forceinline__ T jt_scale(T q, T qdz, T qdy, T qdx)
{
  const T m = fmax(fmax(fabs(q), fabs(qdz)), fmax(fabs(qdy), fabs(qdx)));
  return scalbn(T(1.0), 44 - ilogb(m));
}
// *addr += val on the J tile
template <typename T>
__device__ __forceinline__ void jt_add(T* addr, T val, T scale)
{
  if constexpr (sizeof(T) == 8) {
    const unsigned long long v  = (unsigned long long)__double2ll_rn(val * scale); // two's complement: sums are mod 2^64
    const unsigned           lo = (unsigned)v;
    unsigned*                p  = reinterpret_cast<unsigned*>(addr);
    const unsigned           old = atomicAdd(p, lo);                                // ATOMS.ADD (native)
    const unsigned           hi  = (unsigned)(v >> 32) + ((unsigned)(old + lo) < old ? 1u : 0u);
    if (hi) atomicAdd(p + 1, hi);
  } else {
    (void)scale;
    atomicAdd(addr, val);
  }
}
// what the tile holds at `addr`, as a real number
template <typename T>
__device__ __forceinline__ T jt_read(const T* addr, T inv_scale)
{
  if constexpr (sizeof(T) == 8) return (T)__double_as_longlong(*addr) * inv_scale;
  else return *addr;
}
