// butterfly.cuh -- fp64 complex DFT butterflies of radix 2,3,4,5,6,8,12,16 held in registers.
// Composite radices are Cooley-Tukey products with compile-time internal twiddles
// (48th roots of unity, correctly rounded).
#pragma once
#include "common.cuh"

namespace shb {

// exp(+2 pi i k / 48) = (cos, sin), k = 0..47
constexpr double ROOT48[48][2] = {
    {1.0, 0},   // k=0
    {0.9914448613738104111445575, 0.1305261922200515915484062},   // k=1
    {0.9659258262890682867497432, 0.2588190451025207623488988},   // k=2
    {0.9238795325112867561281832, 0.38268343236508977172846},   // k=3
    {0.8660254037844386467637232, 0.5},   // k=4
    {0.793353340291235164579777, 0.6087614290087206394160975},   // k=5
    {0.7071067811865475244008444, 0.7071067811865475244008444},   // k=6
    {0.6087614290087206394160975, 0.793353340291235164579777},   // k=7
    {0.5, 0.8660254037844386467637232},   // k=8
    {0.38268343236508977172846, 0.9238795325112867561281832},   // k=9
    {0.2588190451025207623488988, 0.9659258262890682867497432},   // k=10
    {0.1305261922200515915484062, 0.9914448613738104111445575},   // k=11
    {0, 1.0},   // k=12
    {-0.1305261922200515915484062, 0.9914448613738104111445575},   // k=13
    {-0.2588190451025207623488988, 0.9659258262890682867497432},   // k=14
    {-0.38268343236508977172846, 0.9238795325112867561281832},   // k=15
    {-0.5, 0.8660254037844386467637232},   // k=16
    {-0.6087614290087206394160975, 0.793353340291235164579777},   // k=17
    {-0.7071067811865475244008444, 0.7071067811865475244008444},   // k=18
    {-0.793353340291235164579777, 0.6087614290087206394160975},   // k=19
    {-0.8660254037844386467637232, 0.5},   // k=20
    {-0.9238795325112867561281832, 0.38268343236508977172846},   // k=21
    {-0.9659258262890682867497432, 0.2588190451025207623488988},   // k=22
    {-0.9914448613738104111445575, 0.1305261922200515915484062},   // k=23
    {-1.0, 0},   // k=24
    {-0.9914448613738104111445575, -0.1305261922200515915484062},   // k=25
    {-0.9659258262890682867497432, -0.2588190451025207623488988},   // k=26
    {-0.9238795325112867561281832, -0.38268343236508977172846},   // k=27
    {-0.8660254037844386467637232, -0.5},   // k=28
    {-0.793353340291235164579777, -0.6087614290087206394160975},   // k=29
    {-0.7071067811865475244008444, -0.7071067811865475244008444},   // k=30
    {-0.6087614290087206394160975, -0.793353340291235164579777},   // k=31
    {-0.5, -0.8660254037844386467637232},   // k=32
    {-0.38268343236508977172846, -0.9238795325112867561281832},   // k=33
    {-0.2588190451025207623488988, -0.9659258262890682867497432},   // k=34
    {-0.1305261922200515915484062, -0.9914448613738104111445575},   // k=35
    {0, -1.0},   // k=36
    {0.1305261922200515915484062, -0.9914448613738104111445575},   // k=37
    {0.2588190451025207623488988, -0.9659258262890682867497432},   // k=38
    {0.38268343236508977172846, -0.9238795325112867561281832},   // k=39
    {0.5, -0.8660254037844386467637232},   // k=40
    {0.6087614290087206394160975, -0.793353340291235164579777},   // k=41
    {0.7071067811865475244008444, -0.7071067811865475244008444},   // k=42
    {0.793353340291235164579777, -0.6087614290087206394160975},   // k=43
    {0.8660254037844386467637232, -0.5},   // k=44
    {0.9238795325112867561281832, -0.38268343236508977172846},   // k=45
    {0.9659258262890682867497432, -0.2588190451025207623488988},   // k=46
    {0.9914448613738104111445575, -0.1305261922200515915484062},   // k=47
};

// multiply by exp(-/+ 2 pi i K / 48)  (forward: minus sign, inverse: plus sign)
template <int K48, bool INV> __device__ __forceinline__ cplx mul_root48(cplx a) {
    constexpr int K = ((K48 % 48) + 48) % 48;
    if constexpr (K == 0) return a;
    else if constexpr (K == 12) return mul_mi<INV>(a);                 // -/+ i
    else if constexpr (K == 24) return make_double2(-a.x, -a.y);
    else if constexpr (K == 36) return mul_mi<!INV>(a);
    else {
        constexpr double c = ROOT48[K][0];
        constexpr double s = INV ? ROOT48[K][1] : -ROOT48[K][1];
        if constexpr (K % 12 == 6) {
            // |c| == |s| == 1/sqrt2 : 2 mul + 2 add
            constexpr double h = 0.70710678118654752440084436210485;
            constexpr double sc = c > 0 ? 1.0 : -1.0, ss = s > 0 ? 1.0 : -1.0;
            return make_double2(h * (sc * a.x - ss * a.y), h * (ss * a.x + sc * a.y));
        } else {
            return make_double2(fma(a.x, c, -a.y * s), fma(a.x, s, a.y * c));
        }
    }
}

template <bool INV> __device__ __forceinline__ void bfly2(cplx &a, cplx &b) {
    cplx t = csub(a, b);
    a = cadd(a, b);
    b = t;
}

template <bool INV> __device__ __forceinline__ void bfly4(cplx &a0, cplx &a1, cplx &a2, cplx &a3) {
    cplx s02 = cadd(a0, a2), d02 = csub(a0, a2);
    cplx s13 = cadd(a1, a3), d13 = mul_mi<INV>(csub(a1, a3));
    a0 = cadd(s02, s13);
    a2 = csub(s02, s13);
    a1 = cadd(d02, d13);
    a3 = csub(d02, d13);
}

template <bool INV> __device__ __forceinline__ void bfly3(cplx &a0, cplx &a1, cplx &a2) {
    const double s60 = 0.86602540378443864676372317075294;
    cplx s = cadd(a1, a2);
    cplx d = csub(a1, a2);
    cplx m = make_double2(fma(-0.5, s.x, a0.x), fma(-0.5, s.y, a0.y));
    cplx r = INV ? make_double2(-s60 * d.y, s60 * d.x) : make_double2(s60 * d.y, -s60 * d.x);
    a0 = cadd(a0, s);
    a1 = cadd(m, r);
    a2 = csub(m, r);
}

template <bool INV> __device__ __forceinline__ void bfly5(cplx &a0, cplx &a1, cplx &a2, cplx &a3, cplx &a4) {
    const double c1 = 0.30901699437494742410229341718282;   // cos(2pi/5)
    const double c2 = -0.80901699437494742410229341718282;  // cos(4pi/5)
    const double s1 = 0.95105651629515357211643933337938;   // sin(2pi/5)
    const double s2 = 0.58778525229247312916870595463907;   // sin(4pi/5)
    cplx p14 = cadd(a1, a4), m14 = csub(a1, a4);
    cplx p23 = cadd(a2, a3), m23 = csub(a2, a3);
    cplx t1 = make_double2(a0.x + c1 * p14.x + c2 * p23.x, a0.y + c1 * p14.y + c2 * p23.y);
    cplx t2 = make_double2(a0.x + c2 * p14.x + c1 * p23.x, a0.y + c2 * p14.y + c1 * p23.y);
    cplx u1 = make_double2(s1 * m14.x + s2 * m23.x, s1 * m14.y + s2 * m23.y);
    cplx u2 = make_double2(s2 * m14.x - s1 * m23.x, s2 * m14.y - s1 * m23.y);
    cplx iu1 = mul_mi<INV>(u1);
    cplx iu2 = mul_mi<INV>(u2);
    a0 = cadd(a0, cadd(p14, p23));
    a1 = cadd(t1, iu1);
    a4 = csub(t1, iu1);
    a2 = cadd(t2, iu2);
    a3 = csub(t2, iu2);
}

// in-register DFT of radix R on v[0..R-1] (natural order in, natural order out)
template <int R, bool INV> struct Bfly;

template <bool INV> struct Bfly<1, INV> { static __device__ __forceinline__ void run(cplx *) {} };
template <bool INV> struct Bfly<2, INV> { static __device__ __forceinline__ void run(cplx *v) { bfly2<INV>(v[0], v[1]); } };
template <bool INV> struct Bfly<3, INV> { static __device__ __forceinline__ void run(cplx *v) { bfly3<INV>(v[0], v[1], v[2]); } };
template <bool INV> struct Bfly<4, INV> { static __device__ __forceinline__ void run(cplx *v) { bfly4<INV>(v[0], v[1], v[2], v[3]); } };
template <bool INV> struct Bfly<5, INV> { static __device__ __forceinline__ void run(cplx *v) { bfly5<INV>(v[0], v[1], v[2], v[3], v[4]); } };

// Cooley-Tukey composite: R = RA*RB.  x[RB*n1+n2] -> X[k1+RA*k2]
//   y[n2][k1] = DFT_RA over n1 ; y *= W_R^{n2 k1} ; X[k1+RA k2] = DFT_RB over n2
template <int RA, int RB, bool INV> struct BflyCT {
    static constexpr int R = RA * RB;
    static __device__ __forceinline__ void run(cplx *v) {
        cplx y[RB][RA];
#pragma unroll
        for (int n2 = 0; n2 < RB; ++n2) {
#pragma unroll
            for (int n1 = 0; n1 < RA; ++n1) y[n2][n1] = v[RB * n1 + n2];
            Bfly<RA, INV>::run(y[n2]);
        }
        twiddle_rows<1>(y);
#pragma unroll
        for (int k1 = 0; k1 < RA; ++k1) {
            cplx z[RB];
#pragma unroll
            for (int n2 = 0; n2 < RB; ++n2) z[n2] = y[n2][k1];
            Bfly<RB, INV>::run(z);
#pragma unroll
            for (int k2 = 0; k2 < RB; ++k2) v[k1 + RA * k2] = z[k2];
        }
    }
    // compile-time unrolled y[n2][k1] *= W_R^{n2*k1}
    template <int N2> static __device__ __forceinline__ void twiddle_rows(cplx (&y)[RB][RA]) {
        if constexpr (N2 < RB) {
            twiddle_cols<N2, 1>(y);
            twiddle_rows<N2 + 1>(y);
        }
    }
    template <int N2, int K1> static __device__ __forceinline__ void twiddle_cols(cplx (&y)[RB][RA]) {
        if constexpr (K1 < RA) {
            y[N2][K1] = mul_root48<(48 / R) * N2 * K1, INV>(y[N2][K1]);
            twiddle_cols<N2, K1 + 1>(y);
        }
    }
};

template <bool INV> struct Bfly<6, INV> { static __device__ __forceinline__ void run(cplx *v) { BflyCT<3, 2, INV>::run(v); } };
template <bool INV> struct Bfly<8, INV> { static __device__ __forceinline__ void run(cplx *v) { BflyCT<4, 2, INV>::run(v); } };
template <bool INV> struct Bfly<12, INV> { static __device__ __forceinline__ void run(cplx *v) { BflyCT<4, 3, INV>::run(v); } };
template <bool INV> struct Bfly<16, INV> { static __device__ __forceinline__ void run(cplx *v) { BflyCT<4, 4, INV>::run(v); } };

}  // namespace shb
