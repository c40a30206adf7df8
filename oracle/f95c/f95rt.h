// f95rt.h -- run-time support for C++ emitted by oracle/f95c/f95c.py (TEST INFRASTRUCTURE ONLY).
//
// The translator turns the reference's own Fortran 95 sources (read where they lie under /root/reference, never
// copied) into C++ statement by statement; this header supplies what the emitted code leans on: arrays with Fortran
// bounds and column-major storage, blank-padded character variables, the handful of intrinsics the reference uses,
// and unformatted direct-access / list-directed I/O on unit numbers.  Nothing here knows anything about the model.
//
// Arithmetic conventions (they are the ones oracle/README.md lists, i.e. gfortran's without -ffast-math):
//   * expressions are evaluated exactly as parenthesised by Fortran's precedence and left-to-right rule;
//   * x**n with an integer n is libgcc's __powidf2 square-and-multiply (x**2 = x*x, x**3 = x*(x*x), x**4 = (x*x)*(x*x));
//   * x**2.0 is the correctly rounded square (gfortran folds it with MPFR);
//   * SUM / MINVAL / MAXVAL / ANY / ALL run sequentially in array-element order;
//   * REAL(x) without a kind is single precision, default-real literals are single precision.
#pragma once
#include <cfloat>
#include <climits>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <string>
#include <sys/stat.h>
#include <time.h>
#include <type_traits>
#include <vector>

namespace f95 {

// ---------------------------------------------------------------- arrays
struct B {  // one dimension's bounds
    long lo, hi;
    B(long l, long h) : lo(l), hi(h) {}
    B(long h) : lo(1), hi(h) {}
};

template <class T, int R>
struct Arr {
    T* d = nullptr;
    long lo_[R], ext_[R], str_[R];
    bool own = false;

    Arr() {
        for (int k = 0; k < R; ++k) lo_[k] = 1, ext_[k] = 0, str_[k] = 0;
    }
    template <class... Bs, class = std::enable_if_t<sizeof...(Bs) == R>>
    explicit Arr(Bs... bs) {
        allocate(bs...);
    }
    Arr(const Arr& o) : d(o.d), own(false) {  // a copy is a view
        for (int k = 0; k < R; ++k) lo_[k] = o.lo_[k], ext_[k] = o.ext_[k], str_[k] = o.str_[k];
    }
    Arr& operator=(const Arr&) = delete;
    ~Arr() {
        if (own) std::free(d);
    }
    void set_bounds(const B* b) {
        long s = 1;
        for (int k = 0; k < R; ++k) {
            lo_[k] = b[k].lo;
            ext_[k] = b[k].hi >= b[k].lo ? b[k].hi - b[k].lo + 1 : 0;
            str_[k] = s;
            s *= ext_[k];
        }
    }
    template <class... Bs>
    void allocate(Bs... bs) {
        static_assert(sizeof...(Bs) == R, "rank");
        B b[R] = {B(bs)...};
        set_bounds(b);
        long n = size();
        // zero-filled, with one last-dimension slab of slack: the reference reads ior4(ipnt, nlay + 1) in write_array
        // (private_mod.f95:2866-2869 with nlay = 1); that read lands in zeros here instead of in the heap
        long slack = (R > 1 ? str_[R - 1] : 0) + 64;
        if (own) std::free(d);
        d = static_cast<T*>(std::calloc(static_cast<size_t>(n + slack), sizeof(T)));
        if (!d) {
            std::fprintf(stderr, "f95rt: out of memory\n");
            std::exit(3);
        }
        own = true;
    }
    void deallocate() {
        if (own) std::free(d);
        d = nullptr;
        own = false;
        for (int k = 0; k < R; ++k) ext_[k] = 0;
    }
    bool allocated() const { return d != nullptr; }
    // explicit-shape dummy argument: same storage, the dummy's own bounds
    template <class... Bs>
    Arr rebound(Bs... bs) const {
        static_assert(sizeof...(Bs) == R, "rank");
        Arr v;
        B b[R] = {B(bs)...};
        v.set_bounds(b);
        v.d = d;
        v.own = false;
        return v;
    }
    long size() const {
        long n = 1;
        for (int k = 0; k < R; ++k) n *= ext_[k];
        return n;
    }
    long lb(int dim) const { return lo_[dim - 1]; }
    long ub(int dim) const { return lo_[dim - 1] + ext_[dim - 1] - 1; }
    long ext(int dim) const { return ext_[dim - 1]; }
    size_t bytes() const { return static_cast<size_t>(size()) * sizeof(T); }
    void fill(T v) {
        long n = size();
        for (long k = 0; k < n; ++k) d[k] = v;
    }
    template <class... I>
    inline T& operator()(I... i) const {
        static_assert(sizeof...(I) == R, "rank");
        long idx[R] = {static_cast<long>(i)...};
        long off = 0;
        for (int k = 0; k < R; ++k) off += (idx[k] - lo_[k]) * str_[k];
        return d[off];
    }
};

// ---------------------------------------------------------------- character variables
inline std::string rtrim(const std::string& s) {
    size_t n = s.size();
    while (n > 0 && s[n - 1] == ' ') --n;
    return s.substr(0, n);
}
struct FStr {  // CHARACTER(len): assignment pads with blanks or truncates
    std::string s;
    explicit FStr(long len) : s(static_cast<size_t>(len), ' ') {}
    FStr(long len, const std::string& v) : s(static_cast<size_t>(len), ' ') { *this = v; }
    FStr& operator=(const std::string& v) {
        size_t n = s.size();
        std::string t = v.substr(0, n);
        t.resize(n, ' ');
        s = t;
        return *this;
    }
    FStr& operator=(const FStr& o) { return *this = o.s; }
    FStr(const FStr&) = default;
    operator const std::string&() const { return s; }
};
inline std::string f_trim(const std::string& a) { return rtrim(a); }
inline int f_len_trim(const std::string& a) { return static_cast<int>(rtrim(a).size()); }
inline std::string f_adjustl(const std::string& a) {
    size_t k = 0;
    while (k < a.size() && a[k] == ' ') ++k;
    std::string t = a.substr(k);
    t.resize(a.size(), ' ');
    return t;
}
inline std::string f_cat(const std::string& a, const std::string& b) { return a + b; }
inline std::string f_substr(const std::string& a, long lo, long hi) {
    if (hi < lo) return std::string();
    return a.substr(static_cast<size_t>(lo - 1), static_cast<size_t>(hi - lo + 1));
}
inline bool str_eq(const std::string& a, const std::string& b) { return rtrim(a) == rtrim(b); }

// comparisons: strings compare blank-padded, everything else as C++ does
template <class A, class C>
inline bool f_eq(const A& a, const C& b) {
    if constexpr (std::is_convertible_v<A, std::string> && std::is_convertible_v<C, std::string>)
        return str_eq(a, b);
    else
        return a == b;
}
template <class A, class C>
inline bool f_ne(const A& a, const C& b) {
    return !f_eq(a, b);
}

// ---------------------------------------------------------------- intrinsics
template <class T>
inline T f_abs(T x) {
    if constexpr (std::is_integral_v<T>)
        return x < 0 ? -x : x;
    else
        return std::fabs(x);
}
template <class A, class C>
inline std::common_type_t<A, C> f_max(A a, C b) {
    using T = std::common_type_t<A, C>;
    return static_cast<T>(a) > static_cast<T>(b) ? static_cast<T>(a) : static_cast<T>(b);
}
template <class A, class C, class... Rest>
inline auto f_max(A a, C b, Rest... r) {
    return f_max(f_max(a, b), r...);
}
template <class A, class C>
inline std::common_type_t<A, C> f_min(A a, C b) {
    using T = std::common_type_t<A, C>;
    return static_cast<T>(a) < static_cast<T>(b) ? static_cast<T>(a) : static_cast<T>(b);
}
template <class A, class C, class... Rest>
inline auto f_min(A a, C b, Rest... r) {
    return f_min(f_min(a, b), r...);
}
inline float f_sqrt(float x) { return std::sqrt(x); }
inline double f_sqrt(double x) { return std::sqrt(x); }
inline float f_cos(float x) { return std::cos(x); }
inline double f_cos(double x) { return std::cos(x); }
inline float f_sin(float x) { return std::sin(x); }
inline double f_sin(double x) { return std::sin(x); }
inline float f_exp(float x) { return std::exp(x); }
inline double f_exp(double x) { return std::exp(x); }
inline int f_floor(double x) { return static_cast<int>(std::floor(x)); }
inline int f_floor(float x) { return static_cast<int>(std::floor(x)); }
inline int f_ceiling(double x) { return static_cast<int>(std::ceil(x)); }
inline int f_ceiling(float x) { return static_cast<int>(std::ceil(x)); }
inline float f_log(float x) { return std::log(x); }
inline double f_log(double x) { return std::log(x); }
inline float f_tanh(float x) { return std::tanh(x); }
inline double f_tanh(double x) { return std::tanh(x); }
inline float f_atan(float x) { return std::atan(x); }
inline double f_atan(double x) { return std::atan(x); }
inline int f_nint(double x) { return static_cast<int>(std::lround(x)); }
inline int f_nint(float x) { return static_cast<int>(std::lroundf(x)); }
template <class A, class C>
inline auto f_mod(A a, C b) {
    if constexpr (std::is_integral_v<A> && std::is_integral_v<C>)
        return a % b;
    else
        return std::fmod(a, b);
}
template <class A, class C>
inline A f_sign(A a, C b) {
    if constexpr (std::is_integral_v<A>)
        return b >= 0 ? f_abs(a) : -f_abs(a);
    else
        return std::copysign(a, static_cast<A>(b));
}
// REAL(x) without a kind: default (single precision) real -- or the real part of a complex, in its own kind
inline float f_real(double x) { return static_cast<float>(x); }
inline float f_real(float x) { return x; }
inline float f_real(int x) { return static_cast<float>(x); }
inline float f_real(long x) { return static_cast<float>(x); }
inline double f_real(const std::complex<double>& z) { return z.real(); }
inline float f_real(const std::complex<float>& z) { return z.real(); }
inline double f_aimag(const std::complex<double>& z) { return z.imag(); }
inline float f_aimag(const std::complex<float>& z) { return z.imag(); }
inline float f_tiny(float) { return FLT_MIN; }
inline double f_tiny(double) { return DBL_MIN; }
inline float f_huge(float) { return FLT_MAX; }
inline double f_huge(double) { return DBL_MAX; }
inline int f_huge(int) { return INT_MAX; }

template <class T>
inline T powi(T x, int m) {  // libgcc __powidf2 / __powisf2
    unsigned n = m < 0 ? 0u - static_cast<unsigned>(m) : static_cast<unsigned>(m);
    T y = (n % 2) ? x : static_cast<T>(1);
    while (n >>= 1) {
        x = x * x;
        if (n % 2) y = y * x;
    }
    return m < 0 ? static_cast<T>(1) / y : y;
}
inline int ipow(int x, int m) {
    int y = 1;
    for (int k = 0; k < m; ++k) y *= x;
    return y;
}
template <class A, class C>
inline auto f_pow(A a, C b) {
    if constexpr (std::is_integral_v<A> && std::is_integral_v<C>)
        return ipow(a, b);
    else if constexpr (std::is_integral_v<C>)
        return powi(a, static_cast<int>(b));
    else {
        using T = std::common_type_t<A, C>;
        T x = static_cast<T>(a), y = static_cast<T>(b);
        if (y == static_cast<T>(2)) return static_cast<T>(x * x);
        return static_cast<T>(std::pow(x, y));
    }
}

// ---------------------------------------------------------------- I/O on unit numbers
struct Unit {
    FILE* f = nullptr;
    long recl = 0;
    bool formatted = false;
    std::string action, name;
};
inline std::map<int, Unit>& units() {
    static std::map<int, Unit> u;
    return u;
}
struct Io {  // the control list of one I/O statement
    int unit = -1, rec = 0, iostat = 0;
    long recl = 0;
    bool has_iostat = false, has_rec = false, list = false;
    std::string file, form = "", access = "sequential", status = "unknown", action = "readwrite", position = "asis", fmt;
};
inline std::string lower(std::string s) {
    for (auto& c : s) c = static_cast<char>(std::tolower(static_cast<unsigned char>(c)));
    return rtrim(s);
}
inline void io_fail(Io& io, const char* what) {
    if (io.has_iostat) {
        io.iostat = 5002;
        return;
    }
    std::fprintf(stderr, "f95rt: %s failed on unit %d (%s)\n", what, io.unit, io.file.c_str());
    std::exit(2);
}
inline bool file_exists(const std::string& p) {
    struct stat st;
    return ::stat(rtrim(p).c_str(), &st) == 0;
}
inline void f_open(Io& io) {
    io.iostat = 0;
    std::string name = rtrim(io.file), st = lower(io.status), ac = lower(io.action), po = lower(io.position);
    auto& tab = units();
    auto it = tab.find(io.unit);
    if (it != tab.end() && it->second.f) {  // OPEN on a connected unit: the old file is closed first
        std::fclose(it->second.f);
        tab.erase(it);
    }
    const char* mode;
    bool exists = file_exists(name);
    if (st == "old" && !exists) return io_fail(io, "open (status='old', no such file)");
    if (st == "replace")
        mode = "w+b";
    else if (po == "append")
        mode = "a+b";
    else if (ac == "read")
        mode = "rb";
    else
        mode = exists ? "r+b" : "w+b";
    FILE* f = std::fopen(name.c_str(), mode);
    if (!f) return io_fail(io, "open");
    Unit u;
    u.f = f;
    u.recl = io.recl;
    u.formatted = lower(io.form) == "formatted" || (io.form.empty() && lower(io.access) == "sequential");
    u.action = ac;
    u.name = name;
    tab[io.unit] = u;
}
inline void f_close(Io& io) {
    io.iostat = 0;
    auto& tab = units();
    auto it = tab.find(io.unit);
    if (it == tab.end()) return;
    if (it->second.f) std::fclose(it->second.f);
    tab.erase(it);
}
inline bool f_opened(int unit) { return units().count(unit) != 0; }
inline std::string f_unit_action(int unit) {
    auto it = units().find(unit);
    if (it == units().end()) return "UNDEFINED";
    std::string a = it->second.action;
    for (auto& c : a) c = static_cast<char>(std::toupper(static_cast<unsigned char>(c)));
    return a;
}
inline Unit* unit_of(Io& io, const char* what) {
    auto it = units().find(io.unit);
    if (it == units().end() || !it->second.f) {
        io_fail(io, what);
        return nullptr;
    }
    return &it->second;
}
// unformatted direct access: the I/O list is laid down contiguously inside record io.rec
struct DirectXfer {
    Io& io;
    Unit* u;
    long off = 0;
    bool writing;
    DirectXfer(Io& i, bool w) : io(i), u(unit_of(i, w ? "write" : "read")), writing(w) {
        io.iostat = 0;
        if (u && std::fseek(u->f, static_cast<long>(io.rec - 1) * u->recl, SEEK_SET) != 0) io_fail(io, "seek");
    }
    void item(void* p, size_t bytes) {
        if (!u || io.iostat) return;
        size_t n = writing ? std::fwrite(p, 1, bytes, u->f) : std::fread(p, 1, bytes, u->f);
        if (n != bytes) io_fail(io, writing ? "write" : "read");
        off += static_cast<long>(bytes);
    }
    template <class T, int R>
    void item(const Arr<T, R>& a) {
        item(a.d, a.bytes());
    }
    template <class T>
    void scalar(T& v) {
        item(&v, sizeof(T));
    }
    void end() {
        if (u && writing) std::fflush(u->f);
    }
};
// list-directed output (unit 6 = stdout, 0 = stderr, else the connected formatted file), one record per statement
struct ListOut {
    Io& io;
    FILE* f = nullptr;
    std::string* internal = nullptr;
    std::string buf;
    explicit ListOut(Io& i) : io(i) {
        io.iostat = 0;
        if (io.unit == 6)
            f = stdout;
        else if (io.unit == 0)
            f = stderr;
        else {
            Unit* u = unit_of(io, "write");
            f = u ? u->f : nullptr;
        }
    }
    ListOut(Io& i, std::string* dst) : io(i), internal(dst) {}
    void put(const std::string& s) { buf += " " + rtrim(s); }
    void put(const char* s) { buf += std::string(" ") + s; }
    void put(double v) {
        char t[64];
        std::snprintf(t, sizeof t, "  %.17g", v);
        buf += t;
    }
    void put(float v) {
        char t[64];
        std::snprintf(t, sizeof t, "  %.9g", static_cast<double>(v));
        buf += t;
    }
    void put(int v) {
        char t[32];
        if (!io.fmt.empty() && io.fmt.find('i') != std::string::npos)
            std::snprintf(t, sizeof t, "%8d", v);  // the one edit descriptor the reference uses: (1i8)
        else
            std::snprintf(t, sizeof t, " %11d", v);
        buf += t;
    }
    void put(long v) { put(static_cast<int>(v)); }
    void put(bool v) { buf += v ? " T" : " F"; }
    void put(const FStr& s) { put(s.s); }
    void end() {
        if (internal) {
            *internal = buf;
            return;
        }
        if (f) {
            std::fprintf(f, "%s\n", buf.c_str());
            std::fflush(f);
        }
    }
};
// list-directed input of numbers from a connected formatted file
struct ListIn {
    Io& io;
    Unit* u;
    explicit ListIn(Io& i) : io(i), u(unit_of(i, "read")) { io.iostat = 0; }
    void get(double& v) {
        if (!u || io.iostat) return;
        if (std::fscanf(u->f, "%lf", &v) != 1) {
            io.iostat = -1;
            if (!io.has_iostat) io_fail(io, "read");
        }
    }
    void get(float& v) {
        double t = 0;
        get(t);
        v = static_cast<float>(t);
    }
    void get(int& v) {
        double t = 0;
        get(t);
        v = static_cast<int>(t);
    }
    void end() {}
};

// ---------------------------------------------------------------- harness knobs (timing runs only)
inline double env_num(const char* name, double compiled) {
    const char* v = std::getenv(name);
    return v ? std::atof(v) : compiled;
}
inline std::string env_str(const char* name, const std::string& compiled) {
    const char* v = std::getenv(name);
    return v ? std::string(v) : compiled;
}
struct TraceLog {
    std::string names;
    std::vector<double> t;
    std::vector<int> id;
};
inline TraceLog& trace_log() {
    static TraceLog l;
    return l;
}
inline void trace_mark(const char* name) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    auto& l = trace_log();
    l.t.push_back(ts.tv_sec + 1e-9 * ts.tv_nsec);
    l.id.push_back(name[0] == 's' && name[1] == 't' && name[2] == 'o' ? 1 : 0);
}
inline void trace_write(const std::string& path) {
    FILE* f = std::fopen(rtrim(path).c_str(), "w");
    if (!f) return;
    auto& l = trace_log();
    for (size_t k = 0; k < l.t.size(); ++k) std::fprintf(f, "%d %.9f\n", l.id[k], l.t[k]);
    std::fclose(f);
}

// ---------------------------------------------------------------- the harness's state dump
// Not part of the reference: at STOP the driver writes every module array in raw form so that a test can compare
// doubles bit for bit (the reference's own records are float32).  File layout, per variable: 32-byte name,
// int32 type (1 = int32, 4 = float32, 8 = float64, 2 = bool), int32 rank, int64 lower bounds[rank], int64 extents[rank], data.
struct Dump {
    FILE* f;
    explicit Dump(const std::string& path) : f(std::fopen(path.c_str(), "wb")) {}
    ~Dump() {
        if (f) std::fclose(f);
    }
    template <class T>
    static int code() {
        if (std::is_same_v<T, double>) return 8;
        if (std::is_same_v<T, float>) return 4;
        if (std::is_same_v<T, bool>) return 2;
        return 1;
    }
    void head(const char* name, int type, int rank, const long* lo, const long* ext) {
        char nm[32] = {0};
        std::strncpy(nm, name, 31);
        std::fwrite(nm, 1, 32, f);
        int32_t h[2] = {type, rank};
        std::fwrite(h, 4, 2, f);
        for (int k = 0; k < rank; ++k) {
            int64_t v = lo[k];
            std::fwrite(&v, 8, 1, f);
        }
        for (int k = 0; k < rank; ++k) {
            int64_t v = ext[k];
            std::fwrite(&v, 8, 1, f);
        }
    }
    template <class T, int R>
    void put(const char* name, const Arr<T, R>& a) {
        if (!f || !a.d) return;
        head(name, code<T>(), R, a.lo_, a.ext_);
        std::fwrite(a.d, sizeof(T), static_cast<size_t>(a.size()), f);
    }
    template <class T>
    void put(const char* name, const T& v) {
        if (!f) return;
        if constexpr (std::is_arithmetic_v<T>) {
            head(name, code<T>(), 0, nullptr, nullptr);
            std::fwrite(&v, sizeof(T), 1, f);
        }
    }
};

}  // namespace f95
