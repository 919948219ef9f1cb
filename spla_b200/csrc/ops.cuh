// ops.cuh -- spla's built-in binary / select ops as device functors (reference src/op.cpp:194-266),
// specialised at compile time for the semirings the BFS / SSSP / PageRank loops use and switchable at
// run time for every other built-in pair. All three value types are 4 bytes.
#pragma once

#include "common.cuh"

#include <cfloat>
#include <climits>
#include <type_traits>

namespace splacu {

    template<typename T> struct TypeCode;
    template<> struct TypeCode<int32_t>  { static constexpr int value = SPLACU_INT; };
    template<> struct TypeCode<uint32_t> { static constexpr int value = SPLACU_UINT; };
    template<> struct TypeCode<float>    { static constexpr int value = SPLACU_FLOAT; };

    template<typename T> __host__ __device__ __forceinline__ T from_bits(uint32_t b) {
        union { uint32_t u; T t; } c; c.u = b; return c.t;
    }
    template<typename T> __host__ __device__ __forceinline__ uint32_t to_bits(T t) {
        union { uint32_t u; T t; } c; c.t = t; return c.u;
    }

    // ---- binary ops -------------------------------------------------------------------------
    // INT arithmetic wraps (computed in uint32); FLOAT uses non-contracting intrinsics so that each
    // product / sum is rounded exactly like the reference's separate std::function calls (no FMA).
    // MIN / MAX follow std::min / std::max: min(a,b) = (b<a)?b:a, max(a,b) = (a<b)?b:a.
    // LOR / LAND return the C++ bool as T, i.e. exactly 0 or 1.

    template<int OP> struct BinStatic {
        static __device__ __forceinline__ float apply(float a, float b) {
            if constexpr (OP == SPLACU_PLUS) return __fadd_rn(a, b);
            else if constexpr (OP == SPLACU_MINUS) return __fsub_rn(a, b);
            else if constexpr (OP == SPLACU_MULT) return __fmul_rn(a, b);
            else if constexpr (OP == SPLACU_DIV) return __fdiv_rn(a, b);
            else if constexpr (OP == SPLACU_MINUS_POW2) { float d = __fsub_rn(a, b); return __fmul_rn(d, d); }
            else if constexpr (OP == SPLACU_FIRST) return a;
            else if constexpr (OP == SPLACU_SECOND) return b;
            else if constexpr (OP == SPLACU_BONE) return 1.0f;
            else if constexpr (OP == SPLACU_MIN) return (b < a) ? b : a;
            else if constexpr (OP == SPLACU_MAX) return (a < b) ? b : a;
            else if constexpr (OP == SPLACU_LOR) return (a != 0.0f || b != 0.0f) ? 1.0f : 0.0f;
            else if constexpr (OP == SPLACU_LAND) return (a != 0.0f && b != 0.0f) ? 1.0f : 0.0f;
            else return 0.0f;// bitwise ops are not defined for FLOAT (rejected on the host)
        }
        static __device__ __forceinline__ uint32_t apply(uint32_t a, uint32_t b) {
            if constexpr (OP == SPLACU_PLUS) return a + b;
            else if constexpr (OP == SPLACU_MINUS) return a - b;
            else if constexpr (OP == SPLACU_MULT) return a * b;
            else if constexpr (OP == SPLACU_DIV) return b ? a / b : 0u;
            else if constexpr (OP == SPLACU_MINUS_POW2) return (a - b) * (a - b);
            else if constexpr (OP == SPLACU_FIRST) return a;
            else if constexpr (OP == SPLACU_SECOND) return b;
            else if constexpr (OP == SPLACU_BONE) return 1u;
            else if constexpr (OP == SPLACU_MIN) return (b < a) ? b : a;
            else if constexpr (OP == SPLACU_MAX) return (a < b) ? b : a;
            else if constexpr (OP == SPLACU_LOR) return (a || b) ? 1u : 0u;
            else if constexpr (OP == SPLACU_LAND) return (a && b) ? 1u : 0u;
            else if constexpr (OP == SPLACU_BOR) return a | b;
            else if constexpr (OP == SPLACU_BAND) return a & b;
            else if constexpr (OP == SPLACU_BXOR) return a ^ b;
            else return 0u;
        }
        static __device__ __forceinline__ int32_t apply(int32_t a, int32_t b) {
            if constexpr (OP == SPLACU_DIV) return (b == 0) ? 0 : ((a == INT_MIN && b == -1) ? INT_MIN : a / b);
            else if constexpr (OP == SPLACU_MIN) return (b < a) ? b : a;
            else if constexpr (OP == SPLACU_MAX) return (a < b) ? b : a;
            else return (int32_t) apply((uint32_t) a, (uint32_t) b);
        }
    };

    template<typename T> __device__ __forceinline__ T bin_dynamic(int op, T a, T b) {
        switch (op) {
            case SPLACU_PLUS: return BinStatic<SPLACU_PLUS>::apply(a, b);
            case SPLACU_MINUS: return BinStatic<SPLACU_MINUS>::apply(a, b);
            case SPLACU_MULT: return BinStatic<SPLACU_MULT>::apply(a, b);
            case SPLACU_DIV: return BinStatic<SPLACU_DIV>::apply(a, b);
            case SPLACU_MINUS_POW2: return BinStatic<SPLACU_MINUS_POW2>::apply(a, b);
            case SPLACU_FIRST: return a;
            case SPLACU_SECOND: return b;
            case SPLACU_BONE: return BinStatic<SPLACU_BONE>::apply(a, b);
            case SPLACU_MIN: return BinStatic<SPLACU_MIN>::apply(a, b);
            case SPLACU_MAX: return BinStatic<SPLACU_MAX>::apply(a, b);
            case SPLACU_LOR: return BinStatic<SPLACU_LOR>::apply(a, b);
            case SPLACU_LAND: return BinStatic<SPLACU_LAND>::apply(a, b);
            case SPLACU_BOR: return BinStatic<SPLACU_BOR>::apply(a, b);
            case SPLACU_BAND: return BinStatic<SPLACU_BAND>::apply(a, b);
            default: return BinStatic<SPLACU_BXOR>::apply(a, b);
        }
    }

    // identity e of an associative + commutative add op: op(e, x) == x for every x the op can produce.
    // (FLOAT PLUS: 0 + (-0) = +0 is the one bit-level exception; FLOAT MIN/MAX use +-inf.)
    template<typename T> __host__ __device__ inline T add_identity(int op) {
        switch (op) {
            case SPLACU_MULT: return T(1);
            case SPLACU_LAND: return T(1);
            case SPLACU_MIN:
                if constexpr (std::is_floating_point_v<T>) return from_bits<T>(0x7f800000u);
                else if constexpr (std::is_signed_v<T>) return T(INT_MAX);
                else return T(UINT_MAX);
            case SPLACU_MAX:
                if constexpr (std::is_floating_point_v<T>) return from_bits<T>(0xff800000u);
                else if constexpr (std::is_signed_v<T>) return T(INT_MIN);
                else return T(0);
            case SPLACU_BAND: return from_bits<T>(0xffffffffu);
            default: return T(0);// PLUS, LOR, BOR, BXOR
        }
    }

    inline bool is_assoc_commutative(int op) {
        switch (op) {
            case SPLACU_PLUS: case SPLACU_MULT: case SPLACU_MIN: case SPLACU_MAX: case SPLACU_LOR:
            case SPLACU_LAND: case SPLACU_BOR: case SPLACU_BAND: case SPLACU_BXOR: return true;
            default: return false;
        }
    }
    inline bool op_valid_for(int dtype, int op) {
        if (op < 0 || op >= SPLACU_BINOP_COUNT) return false;
        if (dtype == SPLACU_FLOAT && (op == SPLACU_BOR || op == SPLACU_BAND || op == SPLACU_BXOR)) return false;
        return true;
    }

    // ---- semirings ---------------------------------------------------------------------------
    // A semiring object is passed to kernels by value; static ones are empty and fold to straight-line code.
    template<typename T, int MUL, int ADD> struct SemiringStatic {
        static constexpr bool is_static = true;
        __device__ __forceinline__ T mult(T a, T b) const { return BinStatic<MUL>::apply(a, b); }
        __device__ __forceinline__ T add(T a, T b) const { return BinStatic<ADD>::apply(a, b); }
        __host__ __device__ __forceinline__ T    identity() const { return add_identity<T>(ADD); }
        __host__ __device__ __forceinline__ int  add_op() const { return ADD; }
        __host__ __device__ __forceinline__ int  mult_op() const { return MUL; }
    };
    template<typename T> struct SemiringDynamic {
        static constexpr bool is_static = false;
        int mul, ad;
        T   ident;
        __device__ __forceinline__ T mult(T a, T b) const { return bin_dynamic<T>(mul, a, b); }
        __device__ __forceinline__ T add(T a, T b) const { return bin_dynamic<T>(ad, a, b); }
        __host__ __device__ __forceinline__ T    identity() const { return ident; }
        __host__ __device__ __forceinline__ int  add_op() const { return ad; }
        __host__ __device__ __forceinline__ int  mult_op() const { return mul; }
    };

    // ---- select ops -------------------------------------------------------------------------
    // encoded as a 4-bit set over the classes {x<0, x==0, x>0, unordered(NaN)}
    struct Select {
        uint32_t classes;
        bool     reads_mask;// false for ALWAYS / NEVER: the mask array is never touched
        template<typename T> __device__ __forceinline__ bool test(T x) const {
            uint32_t c;
            if (x < T(0)) c = 1u;
            else if (x == T(0)) c = 2u;
            else if (x > T(0)) c = 4u;
            else c = 8u;
            return (classes & c) != 0u;
        }
    };
    inline Select make_select(int op) {
        switch (op) {
            case SPLACU_EQZERO: return {2u, true};
            case SPLACU_NQZERO: return {1u | 4u | 8u, true};
            case SPLACU_GTZERO: return {4u, true};
            case SPLACU_GEZERO: return {2u | 4u, true};
            case SPLACU_LTZERO: return {1u, true};
            case SPLACU_LEZERO: return {1u | 2u, true};
            case SPLACU_ALWAYS: return {15u, false};
            default: return {0u, false};// NEVER
        }
    }

    // value comparison `a != b` in T (FLOAT: NaN != NaN, -0 == +0), as the reference compares T values
    template<typename T> __device__ __forceinline__ bool value_neq(T a, T b) { return a != b; }

    // ---- atomics for the push accumulator ------------------------------------------------------
    template<typename T> __device__ __forceinline__ void atomic_cas_combine(int op, T* addr, T val) {
        uint32_t* a   = reinterpret_cast<uint32_t*>(addr);
        uint32_t  old = *a, assumed;
        do {
            assumed   = old;
            T updated = bin_dynamic<T>(op, from_bits<T>(assumed), val);
            if (to_bits(updated) == assumed) return;
            old = atomicCAS(a, assumed, to_bits(updated));
        } while (old != assumed);
    }

    // acc[j] = add(acc[j], val) for an associative + commutative add with acc pre-set to the identity
    // `cur` is a (possibly stale) value of *addr the caller already loaded: it only serves to skip atomics that cannot
    // change the accumulator (idempotent ops); a stale value costs at most one redundant atomic.
    template<typename T> __device__ __forceinline__ void atomic_combine(int op, T* addr, T val, T cur) {
        if constexpr (std::is_floating_point_v<T>) {// float
            float* a = reinterpret_cast<float*>(addr);
            switch (op) {
                case SPLACU_PLUS: atomicAdd(a, val); return;
                case SPLACU_MIN:
                    if (!(val < cur)) return;// idempotent: skip when no change (stale read only costs an atomic)
                    // ordered-int trick, keyed on the SIGN BIT (so that -0.0 takes the negative branch)
                    if (__float_as_int(val) >= 0) atomicMin(reinterpret_cast<int*>(a), __float_as_int(val));
                    else atomicMax(reinterpret_cast<unsigned int*>(a), __float_as_uint(val));
                    return;
                case SPLACU_MAX:
                    if (!(cur < val)) return;
                    if (__float_as_int(val) >= 0) atomicMax(reinterpret_cast<int*>(a), __float_as_int(val));
                    else atomicMin(reinterpret_cast<unsigned int*>(a), __float_as_uint(val));
                    return;
                case SPLACU_LOR:
                    if (val != 0.0f && cur == 0.0f) atomicExch(a, 1.0f);
                    return;
                case SPLACU_LAND:
                    if (val == 0.0f && cur != 0.0f) atomicExch(a, 0.0f);
                    return;
                default: atomic_cas_combine<T>(op, addr, val); return;// MULT
            }
        } else if constexpr (std::is_signed_v<T>) {// int32
            int* a = reinterpret_cast<int*>(addr);
            switch (op) {
                case SPLACU_PLUS: atomicAdd(a, val); return;
                case SPLACU_MIN: if (val < cur) atomicMin(a, val); return;
                case SPLACU_MAX: if (val > cur) atomicMax(a, val); return;
                case SPLACU_BOR: if ((cur | val) != cur) atomicOr(a, val); return;
                case SPLACU_BAND: if ((cur & val) != cur) atomicAnd(a, val); return;
                case SPLACU_BXOR: atomicXor(a, val); return;
                case SPLACU_LOR: if (val != 0 && cur == 0) atomicExch(a, 1); return;
                case SPLACU_LAND: if (val == 0 && cur != 0) atomicExch(a, 0); return;
                default: atomic_cas_combine<T>(op, addr, val); return;// MULT
            }
        } else {// uint32
            unsigned int* a = reinterpret_cast<unsigned int*>(addr);
            switch (op) {
                case SPLACU_PLUS: atomicAdd(a, val); return;
                case SPLACU_MIN: if (val < cur) atomicMin(a, val); return;
                case SPLACU_MAX: if (val > cur) atomicMax(a, val); return;
                case SPLACU_BOR: if ((cur | val) != cur) atomicOr(a, val); return;
                case SPLACU_BAND: if ((cur & val) != cur) atomicAnd(a, val); return;
                case SPLACU_BXOR: atomicXor(a, val); return;
                case SPLACU_LOR: if (val != 0u && cur == 0u) atomicExch(a, 1u); return;
                case SPLACU_LAND: if (val == 0u && cur != 0u) atomicExch(a, 0u); return;
                default: atomic_cas_combine<T>(op, addr, val); return;// MULT
            }
        }
    }

    // ---- host-side dispatch helpers ---------------------------------------------------------
    // Calls f(T{}) with the C++ type for a dtype code.
    template<typename F> inline int dispatch_dtype(int dtype, F&& f) {
        switch (dtype) {
            case SPLACU_INT: return f(int32_t{});
            case SPLACU_UINT: return f(uint32_t{});
            case SPLACU_FLOAT: return f(float{});
            default: set_error("unknown dtype %d", dtype); return SPLACU_E_INVALID;
        }
    }

    // Calls f(semiring) with a compile-time specialised semiring for the pairs the graph algorithms use
    // (reference src/algorithm.cpp:97-99 BAND/BOR, README LAND/LOR, :208-210 PLUS/MIN, :312 MULT/PLUS)
    // and a run-time switched one for every other built-in pair.
    template<typename T, typename F> inline int dispatch_semiring(int op_mult, int op_add, F&& f) {
        if (op_mult == SPLACU_MULT && op_add == SPLACU_PLUS) return f(SemiringStatic<T, SPLACU_MULT, SPLACU_PLUS>{});
        if (op_mult == SPLACU_PLUS && op_add == SPLACU_MIN) return f(SemiringStatic<T, SPLACU_PLUS, SPLACU_MIN>{});
        if (op_mult == SPLACU_LAND && op_add == SPLACU_LOR) return f(SemiringStatic<T, SPLACU_LAND, SPLACU_LOR>{});
        if constexpr (!std::is_floating_point_v<T>) {
            if (op_mult == SPLACU_BAND && op_add == SPLACU_BOR) return f(SemiringStatic<T, SPLACU_BAND, SPLACU_BOR>{});
        }
        SemiringDynamic<T> s;
        s.mul   = op_mult;
        s.ad    = op_add;
        s.ident = add_identity<T>(op_add);
        return f(s);
    }

}// namespace splacu
