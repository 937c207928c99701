// ORACLE / TEST INFRASTRUCTURE — not product code.
//
// GLSL 4.30 types and built-ins as C++20, just enough to compile the reference's compute shader
// (/root/reference/shaders/compute/falling_sand.glsl and its includes) with g++ so that the reference's OWN shader
// text runs on the CPU (oracle/build_ref.py makes the translation unit; the result lives in oracle/_ref/).
// Nothing of the reference is copied here: this file only supplies what a GLSL compiler has built in —
//   * vecN / ivecN / uvecN with component views (.x .r .xy .rgb ...), constructors that flatten their arguments,
//     component-wise operators with GLSL's implicit int -> uint -> float promotions, == / != yielding bool;
//   * the built-in functions the shader calls (floor ceil fract sign mod sin cos tan abs sqrt pow step min max clamp
//     mix dot distance), all evaluated in f32 (float literals get an `f` suffix in the translation, -ffp-contract=off);
//   * sampler2D / image2D over plain RGBA32F host arrays, texelFetch / imageStore, gl_GlobalInvocationID.
// Everything is in namespace glsl; the translated shader is placed in the same namespace so its unqualified
// calls (sin, pow, abs, ...) resolve here and never to <cmath>'s double overloads.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <type_traits>

namespace glsl {

typedef unsigned int uint;

template <class T, int N> struct vec;

// A view of some components of a parent vector's storage (member of the parent's anonymous union).
template <class T, int P, int... I> struct swz {
    T d[P];
    static constexpr int size = (int)sizeof...(I);
    T get(int i) const { constexpr int idx[] = {I...}; return d[idx[i]]; }
    template <class V> swz& operator=(const V& v);   // defined after vec (distinct indices only)
};

// ---- traits -------------------------------------------------------------------------------------------------
template <class A> struct vt { static constexpr bool is_vec = false; };
template <class T, int N> struct vt<vec<T, N>> {
    static constexpr bool is_vec = true; static constexpr int size = N; using elem = T;
    static T get(const vec<T, N>& v, int i) { return v.d[i]; }
};
template <class T, int P, int... I> struct vt<swz<T, P, I...>> {
    static constexpr bool is_vec = true; static constexpr int size = (int)sizeof...(I); using elem = T;
    static T get(const swz<T, P, I...>& v, int i) { return v.get(i); }
};
template <class A> concept vecish = vt<A>::is_vec;
template <class A> concept scalar = std::is_arithmetic_v<A>;
template <class A> concept operand = vecish<A> || scalar<A>;

template <class A> struct elem_of { using type = A; };
template <vecish A> struct elem_of<A> { using type = typename vt<A>::elem; };
template <class A> using elem_t = typename elem_of<A>::type;
template <class A> constexpr int comps() { if constexpr (vecish<A>) return vt<A>::size; else return 1; }
template <class A> auto comp(const A& a, int i) { if constexpr (vecish<A>) return vt<A>::get(a, i); else return a; }
// width of the result of a component-wise operation on A and B (a scalar broadcasts)
template <class A, class B> constexpr int width() {
    constexpr int a = comps<A>(), b = comps<B>();
    static_assert(!(vecish<A> && vecish<B>) || a == b, "component counts differ");
    return vecish<A> ? a : b;
}
template <class A, class B> using promote_t = std::common_type_t<elem_t<A>, elem_t<B>>;   // int -> uint -> float

// ---- storage ------------------------------------------------------------------------------------------------
template <class T> struct storage2 {
    union {
        T d[2];
        struct { T x, y; };
        struct { T r, g; };
        swz<T, 2, 0, 1> xy; swz<T, 2, 1, 0> yx; swz<T, 2, 0, 1> rg;
    };
};
template <class T> struct storage3 {
    union {
        T d[3];
        struct { T x, y, z; };
        struct { T r, g, b; };
        swz<T, 3, 0, 1> xy; swz<T, 3, 1, 2> yz; swz<T, 3, 0, 2> xz; swz<T, 3, 0, 1> rg;
        swz<T, 3, 0, 1, 2> xyz; swz<T, 3, 0, 1, 2> rgb;
    };
};
template <class T> struct storage4 {
    union {
        T d[4];
        struct { T x, y, z, w; };
        struct { T r, g, b, a; };
        swz<T, 4, 0, 1> xy; swz<T, 4, 2, 3> zw; swz<T, 4, 1, 2> yz; swz<T, 4, 0, 1> rg;
        swz<T, 4, 0, 1, 2> xyz; swz<T, 4, 0, 1, 2> rgb; swz<T, 4, 1, 2, 3> yzw;
        swz<T, 4, 0, 1, 2, 3> xyzw; swz<T, 4, 0, 1, 2, 3> rgba;
    };
};
template <class T, int N> struct storage_of;
template <class T> struct storage_of<T, 2> { using type = storage2<T>; };
template <class T> struct storage_of<T, 3> { using type = storage3<T>; };
template <class T> struct storage_of<T, 4> { using type = storage4<T>; };

// GLSL converts int -> uint -> float implicitly; everything else needs a constructor call
template <class From, class To> constexpr bool implicit_ok =
    std::is_same_v<From, To> || (std::is_same_v<To, float> && std::is_integral_v<From>) ||
    (std::is_same_v<To, uint> && std::is_same_v<From, int>);

template <class T, int N> struct vec : storage_of<T, N>::type {
    using storage_of<T, N>::type::d;
    vec() = default;
    // one argument: a scalar splats, a vector of >= N components converts (and truncates)
    template <operand A> requires (comps<A>() == 1 || comps<A>() >= N)
    explicit(!(vecish<A> && comps<A>() == N && implicit_ok<elem_t<A>, T>)) vec(const A& a) {
        for (int i = 0; i < N; ++i) d[i] = static_cast<T>(comp(a, comps<A>() == 1 ? 0 : i));
    }
    // several arguments: their components are laid out one after the other and must add up to N
    template <operand A, operand B, operand... C> requires (comps<A>() + comps<B>() + (comps<C>() + ... + 0) == N)
    vec(const A& a, const B& b, const C&... c) {
        int n = 0;
        put(n, a); put(n, b); (put(n, c), ...);
    }
    T& operator[](int i) { return d[i]; }
    const T& operator[](int i) const { return d[i]; }
    int length() const { return N; }

    template <operand B> vec& operator+=(const B& b) { return *this = vec(*this + b); }
    template <operand B> vec& operator-=(const B& b) { return *this = vec(*this - b); }
    template <operand B> vec& operator*=(const B& b) { return *this = vec(*this * b); }
    template <operand B> vec& operator/=(const B& b) { return *this = vec(*this / b); }
    template <operand B> vec& operator%=(const B& b) { return *this = vec(*this % b); }
    template <operand B> vec& operator&=(const B& b) { return *this = vec(*this & b); }
    template <operand B> vec& operator|=(const B& b) { return *this = vec(*this | b); }
    template <operand B> vec& operator^=(const B& b) { return *this = vec(*this ^ b); }
    template <operand B> vec& operator<<=(const B& b) { return *this = vec(*this << b); }
    template <operand B> vec& operator>>=(const B& b) { return *this = vec(*this >> b); }

private:
    template <class A> void put(int& n, const A& a) {
        for (int i = 0; i < comps<A>(); ++i) d[n++] = static_cast<T>(comp(a, i));
    }
};

template <class T, int P, int... I> template <class V> swz<T, P, I...>& swz<T, P, I...>::operator=(const V& v) {
    constexpr int idx[] = {I...};
    const vec<T, (int)sizeof...(I)> tmp(v);
    for (int i = 0; i < (int)sizeof...(I); ++i) d[idx[i]] = tmp.d[i];
    return *this;
}

using vec2 = vec<float, 2>;  using vec3 = vec<float, 3>;  using vec4 = vec<float, 4>;
using ivec2 = vec<int, 2>;   using ivec3 = vec<int, 3>;   using ivec4 = vec<int, 4>;
using uvec2 = vec<uint, 2>;  using uvec3 = vec<uint, 3>;  using uvec4 = vec<uint, 4>;
using bvec2 = vec<bool, 2>;  using bvec3 = vec<bool, 3>;  using bvec4 = vec<bool, 4>;

// ---- component-wise operators -------------------------------------------------------------------------------
template <class R, class A, class B, class F> auto zip(const A& a, const B& b, F f) {
    constexpr int N = width<A, B>();
    vec<R, N> out;
    for (int i = 0; i < N; ++i) out.d[i] = f(static_cast<R>(comp(a, i)), static_cast<R>(comp(b, i)));
    return out;
}
#define GLSL_BINOP(op)                                                                                      \
    template <operand A, operand B> requires (vecish<A> || vecish<B>)                                       \
    auto operator op(const A& a, const B& b) {                                                              \
        using R = promote_t<A, B>;                                                                          \
        return zip<R>(a, b, [](R x, R y) { return static_cast<R>(x op y); });                               \
    }
GLSL_BINOP(+) GLSL_BINOP(-) GLSL_BINOP(*) GLSL_BINOP(/)
#undef GLSL_BINOP
#define GLSL_INTOP(op)                                                                                      \
    template <operand A, operand B> requires (vecish<A> || vecish<B>) && std::is_integral_v<promote_t<A, B>> \
    auto operator op(const A& a, const B& b) {                                                              \
        using R = promote_t<A, B>;                                                                          \
        return zip<R>(a, b, [](R x, R y) { return static_cast<R>(x op y); });                               \
    }
GLSL_INTOP(%) GLSL_INTOP(&) GLSL_INTOP(|) GLSL_INTOP(^)
#undef GLSL_INTOP
// shifts keep the type of the left operand
#define GLSL_SHIFT(op)                                                                                      \
    template <vecish A, operand B> requires std::is_integral_v<elem_t<A>> && std::is_integral_v<elem_t<B>>  \
    auto operator op(const A& a, const B& b) {                                                              \
        using R = elem_t<A>;                                                                                \
        vec<R, comps<A>()> out;                                                                             \
        for (int i = 0; i < comps<A>(); ++i) out.d[i] = static_cast<R>(comp(a, i) op comp(b, comps<B>() == 1 ? 0 : i)); \
        return out;                                                                                         \
    }
GLSL_SHIFT(<<) GLSL_SHIFT(>>)
#undef GLSL_SHIFT
template <vecish A> auto operator-(const A& a) {
    vec<elem_t<A>, comps<A>()> out;
    for (int i = 0; i < comps<A>(); ++i) out.d[i] = -comp(a, i);
    return out;
}
// GLSL: == and != on vectors compare the whole value and give ONE bool
template <vecish A, vecish B> requires (comps<A>() == comps<B>())
bool operator==(const A& a, const B& b) {
    using R = promote_t<A, B>;
    for (int i = 0; i < comps<A>(); ++i)
        if (!(static_cast<R>(comp(a, i)) == static_cast<R>(comp(b, i)))) return false;
    return true;
}
template <vecish A, vecish B> requires (comps<A>() == comps<B>())
bool operator!=(const A& a, const B& b) { return !(a == b); }

// ---- built-in functions (f32) -------------------------------------------------------------------------------
template <class A, class F> auto map1f(const A& a, F f) {          // A promoted to float, like a genType argument
    if constexpr (vecish<A>) {
        vec<float, comps<A>()> out;
        for (int i = 0; i < comps<A>(); ++i) out.d[i] = f(static_cast<float>(comp(a, i)));
        return out;
    } else {
        return f(static_cast<float>(a));
    }
}
template <class A, class B, class F> auto map2f(const A& a, const B& b, F f) {
    if constexpr (vecish<A> || vecish<B>) {
        return zip<float>(a, b, f);
    } else {
        return f(static_cast<float>(a), static_cast<float>(b));
    }
}
template <operand A> auto floor(const A& a) { return map1f(a, [](float x) { return std::floor(x); }); }
template <operand A> auto fract(const A& a) { return map1f(a, [](float x) { return x - std::floor(x); }); }
template <operand A> auto sin(const A& a) { return map1f(a, [](float x) { return std::sin(x); }); }
template <operand A> auto cos(const A& a) { return map1f(a, [](float x) { return std::cos(x); }); }
template <operand A> auto tan(const A& a) { return map1f(a, [](float x) { return std::tan(x); }); }
template <operand A> auto sqrt(const A& a) { return map1f(a, [](float x) { return std::sqrt(x); }); }
template <operand A> auto abs(const A& a) {
    if constexpr (vecish<A>) {
        vec<elem_t<A>, comps<A>()> out;
        for (int i = 0; i < comps<A>(); ++i) { const auto x = comp(a, i); out.d[i] = x < 0 ? -x : x; }
        return out;
    } else {
        return a < 0 ? -a : a;
    }
}
template <operand A> auto ceil(const A& a) { return map1f(a, [](float x) { return std::ceil(x); }); }
template <operand A> auto sign(const A& a) {
    if constexpr (vecish<A>) {
        vec<elem_t<A>, comps<A>()> out;
        for (int i = 0; i < comps<A>(); ++i) { const auto x = comp(a, i); out.d[i] = x > 0 ? 1 : (x < 0 ? -1 : 0); }
        return out;
    } else {
        return static_cast<A>(a > 0 ? 1 : (a < 0 ? -1 : 0));
    }
}
// mod(x, y) = x - y * floor(x / y)   (GLSL 4.30 spec, 8.3)
template <operand A, operand B> auto mod(const A& a, const B& b) { return map2f(a, b, [](float x, float y) { return x - y * std::floor(x / y); }); }
template <operand A, operand B> auto pow(const A& a, const B& b) { return map2f(a, b, [](float x, float y) { return std::pow(x, y); }); }
template <operand A, operand B> auto step(const A& edge, const B& x) { return map2f(edge, x, [](float e, float v) { return v < e ? 0.0f : 1.0f; }); }
template <operand A, operand B> auto max(const A& a, const B& b) {
    using R = promote_t<A, B>;
    if constexpr (vecish<A> || vecish<B>) return zip<R>(a, b, [](R x, R y) { return x < y ? y : x; });
    else return static_cast<R>(a) < static_cast<R>(b) ? static_cast<R>(b) : static_cast<R>(a);
}
template <operand A, operand B> auto min(const A& a, const B& b) {
    using R = promote_t<A, B>;
    if constexpr (vecish<A> || vecish<B>) return zip<R>(a, b, [](R x, R y) { return y < x ? y : x; });
    else return static_cast<R>(b) < static_cast<R>(a) ? static_cast<R>(b) : static_cast<R>(a);
}
template <operand A, operand B, operand C> auto clamp(const A& x, const B& lo, const C& hi) { return min(max(x, lo), hi); }
// mix(x, y, a) = x * (1 - a) + y * a   (GLSL 4.30 spec, 8.3)
template <operand A, operand B, operand C> auto mix(const A& x, const B& y, const C& a) { return x * (1.0f - a) + y * a; }
template <operand A, operand B> float dot(const A& a, const B& b) {
    float s = static_cast<float>(comp(a, 0)) * static_cast<float>(comp(b, 0));
    for (int i = 1; i < width<A, B>(); ++i) s = s + static_cast<float>(comp(a, i)) * static_cast<float>(comp(b, i));
    return s;
}
template <operand A> float length(const A& a) { return std::sqrt(dot(a, a)); }
template <operand A, operand B> float distance(const A& a, const B& b) { return length(a - b); }

struct mat2 {            // column-major like GLSL; only what rotatePoint needs
    float m[4];
    mat2(float a, float b, float c, float d) : m{a, b, c, d} {}
};
inline vec2 operator*(const mat2& M, const vec2& v) { return vec2(M.m[0] * v.x + M.m[2] * v.y, M.m[1] * v.x + M.m[3] * v.y); }

// ---- arrays: `T[N] name`, `.length()` -------------------------------------------------------------------------
template <class T, int N> struct glsl_array {
    T v[N];
    T& operator[](int i) { return v[i]; }
    const T& operator[](int i) const { return v[i]; }
    int length() const { return N; }
};

// ---- textures and images: RGBA32F host arrays, row-major, y down ----------------------------------------------
struct sampler2D { const vec4* texels = nullptr; int w = 0, h = 0; };
struct image2D { vec4* texels = nullptr; int w = 0, h = 0; };
struct uimage2D { uvec4* texels = nullptr; int w = 0, h = 0; };
inline vec4 texelFetch(const sampler2D& s, const ivec2& p, int /*lod*/) {
    if (p.x < 0 || p.y < 0 || p.x >= s.w || p.y >= s.h) std::abort();      // undefined in GL; the shader never does it
    return s.texels[(size_t)p.y * s.w + p.x];
}
template <vecish V> void imageStore(const image2D& im, const ivec2& p, const V& v) {
    if (!im.texels) return;                                                 // image not bound
    if (p.x < 0 || p.y < 0 || p.x >= im.w || p.y >= im.h) return;           // GL discards out-of-range stores
    im.texels[(size_t)p.y * im.w + p.x] = vec4(v);
}
inline vec4 imageLoad(const image2D& im, const ivec2& p) { return im.texels[(size_t)p.y * im.w + p.x]; }

inline thread_local uvec3 gl_GlobalInvocationID;

}  // namespace glsl
