// jit.cu -- model front-end: a declarative spec of an Unfold kernel becomes a device functor at run time.
//
// In modppl a model is written with the `dyngen!` proc-macro (modppl-macros/src/lib.rs:20-114): ordinary Rust with `%=` sample
// statements (`sample_at`, dyngenfn.rs:115-141), wrapped in a DynUnfold (dynunfold.rs:41-100).  The restricted vectorisable
// form of that -- fixed-shape state, one sample statement per state component at t = 0 and one per step, observation terms from
// the built-in log-densities -- is described here by a small JSON document whose expressions are C++ expressions over the
// previous state x[i], the observation y[i], the time index t and the named parameters:
//
//   { "name": "tracker", "state_dim": 4, "obs_dim": 2,
//     "params":  {"q": 0.1, "r": 0.5, "x0": 1.0},
//     "init":    [{"dist": "normal", "args": ["0", "x0"]}, ... one per state component ...],           // x_d ~ dist        (t = 0)
//     "step":    [{"dist": "normal", "args": ["x[0] + x[2]", "q"]}, ...],                             // x'_d ~ dist(x)     (t > 0)
//     "observe": [{"dist": "normal", "value": "y[0]", "args": ["x[0]", "r"]}, ...] }                  // weight += logpdf(value; args(x'))
//
// sample dists: normal(mean, std), uniform(a, b), delta(value); {"add_to": "x[0]"} turns a sample into an increment
// (`r += normal(0, .1)`, tests/dyngenfns/unfold.rs:22-25).  observe dists: normal(mean, std), uniform(a, b), bernoulli(p),
// mvnormal2 {"value": [..2], "mean": [..2], "cov": [..4]} (tests/dyngenfns/unfold.rs:27-31), expr {"value": log-density}.
// The spec is turned into `struct JitModel<Real>` with the kernel(t, stream, x, obs) member every built-in functor of
// models.cuh has, and NVRTC compiles pf_extend_kernel<JitModel<Real>, ...> from the SAME headers the library was built from
// (embedded at build time) -- the model runs through exactly the kernels of the built-in ones: no rebuild of the library, no
// second code path.  The emitted arithmetic follows the built-in functors' conventions (z * std + mean; normal observation terms
// pooled into one sum of squares with hoisted 1/std and constant), so a spec of lgssm4 reproduces Lgssm4<Real> bit for bit.
#include <dlfcn.h>
#include <nvrtc.h>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <vector>
#include "engine.h"

namespace mpl {

// ---- the kernel headers, embedded at build time (Makefile: lib/embedded_headers.inc) ---------------------------------------------
struct EmbeddedHeader { const char* name; const char* text; };
static const EmbeddedHeader kHeaders[] = {
#include "../lib/embedded_headers.inc"
};

// ---- a minimal JSON reader (objects, arrays, strings, numbers, true/false/null) ---------------------------------------------
struct Json {
    enum Kind { NUL, NUM, STR, ARR, OBJ, BOOL } kind = NUL;
    double num = 0.;
    bool b = false;
    std::string str;
    std::vector<Json> arr;
    std::vector<std::pair<std::string, Json>> obj;   // insertion order kept (parameter order)
    const Json* get(const std::string& k) const { for (auto& kv : obj) if (kv.first == k) return &kv.second; return nullptr; }
};
struct JsonReader {
    const char* p; const char* end; std::string err;
    void ws() { while (p < end && std::isspace((unsigned char)*p)) ++p; }
    bool fail_(const std::string& m) { if (err.empty()) err = m; return false; }
    bool value(Json& out) {
        ws();
        if (p >= end) return fail_("unexpected end of spec");
        if (*p == '{') {
            out.kind = Json::OBJ; ++p; ws();
            if (p < end && *p == '}') { ++p; return true; }
            for (;;) {
                Json k;
                ws();
                if (p >= end || *p != '"' || !string(k.str)) return fail_("object key expected");
                ws();
                if (p >= end || *p != ':') return fail_("':' expected");
                ++p;
                Json v;
                if (!value(v)) return false;
                out.obj.emplace_back(k.str, std::move(v));
                ws();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == '}') { ++p; return true; }
                return fail_("',' or '}' expected");
            }
        }
        if (*p == '[') {
            out.kind = Json::ARR; ++p; ws();
            if (p < end && *p == ']') { ++p; return true; }
            for (;;) {
                Json v;
                if (!value(v)) return false;
                out.arr.push_back(std::move(v));
                ws();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == ']') { ++p; return true; }
                return fail_("',' or ']' expected");
            }
        }
        if (*p == '"') { out.kind = Json::STR; return string(out.str); }
        if (!std::strncmp(p, "true", 4)) { out.kind = Json::BOOL; out.b = true; p += 4; return true; }
        if (!std::strncmp(p, "false", 5)) { out.kind = Json::BOOL; out.b = false; p += 5; return true; }
        if (!std::strncmp(p, "null", 4)) { out.kind = Json::NUL; p += 4; return true; }
        char* e = nullptr;
        out.num = std::strtod(p, &e);
        if (e == p) return fail_("value expected");
        out.kind = Json::NUM; p = e;
        return true;
    }
    bool string(std::string& out) {
        ++p;
        while (p < end && *p != '"') {
            if (*p == '\\' && p + 1 < end) { ++p; out.push_back(*p == 'n' ? '\n' : *p == 't' ? '\t' : *p); ++p; }
            else out.push_back(*p++);
        }
        if (p >= end) return fail_("unterminated string");
        ++p;
        return true;
    }
};

static std::string num_lit(double v) {   // a double literal that round-trips
    char buf[64];
    std::snprintf(buf, sizeof buf, "%.17g", v);
    std::string s(buf);
    if (s.find_first_of(".eEni") == std::string::npos) s += ".";
    return s;
}
// expression of a spec: a JSON string (C++ expression) or a JSON number
static bool expr_of(const Json& j, std::string& out) {
    if (j.kind == Json::STR) { out = j.str; return !out.empty(); }
    if (j.kind == Json::NUM) { out = num_lit(j.num); return true; }
    return false;
}
static bool is_zero_literal(const std::string& e) {
    char* end = nullptr;
    const double v = std::strtod(e.c_str(), &end);
    while (end && *end && std::isspace((unsigned char)*end)) ++end;
    return end && end != e.c_str() && *end == 0 && v == 0.;
}
static bool safe_expr(const std::string& e) {   // expressions are C++ *expressions*: no statements, no preprocessor, no strings
    for (char c : e) if (c == ';' || c == '{' || c == '}' || c == '#' || c == '"' || c == '\'' || c == '\\' || c == '\n') return false;
    return true;
}

// ---- spec -> source --------------------------------------------------------------------------------------------------------------
struct JitSpec {
    std::string name;
    int D = 0, nobs = 0;
    std::vector<std::pair<std::string, double>> params;
    std::string source[2];             // generated translation unit per dtype (MPL_F32, MPL_F64)
    std::vector<double> derived[2];    // constants evaluated on the host in the kernel's precision (1/std, pooled normalisers)
};

struct SampleStmt { std::string dist, a0, a1, add_to; };

static int parse_samples(const Json* arr, int D, const char* what, std::vector<SampleStmt>& out, std::string& err) {
    if (!arr || arr->kind != Json::ARR || (int)arr->arr.size() != D) { err = std::string("\"") + what + "\": one sample statement per state component expected"; return -1; }
    for (const Json& s : arr->arr) {
        SampleStmt st;
        const Json* d = s.get("dist");
        if (s.kind != Json::OBJ || !d || d->kind != Json::STR) { err = std::string(what) + ": {\"dist\": ..., \"args\": [...]} expected"; return -1; }
        st.dist = d->str;
        const Json* args = s.get("args");
        const size_t need = st.dist == "delta" ? 1 : 2;
        if ((st.dist != "normal" && st.dist != "uniform" && st.dist != "delta") || !args || args->kind != Json::ARR || args->arr.size() != need ||
            !expr_of(args->arr[0], st.a0) || (need == 2 && !expr_of(args->arr[1], st.a1))) {
            err = std::string(what) + ": sample dists are normal(mean, std), uniform(a, b), delta(value)"; return -1;
        }
        if (const Json* add = s.get("add_to")) { if (!expr_of(*add, st.add_to)) { err = "add_to: expression expected"; return -1; } }
        if (!safe_expr(st.a0) || !safe_expr(st.a1) || !safe_expr(st.add_to)) { err = "expressions must be plain C++ expressions"; return -1; }
        out.push_back(st);
    }
    return 0;
}

// the sample statements of one branch (t == 0 or t > 0): draws laid out like the built-in functors do -- all normals of the
// branch from draw_normals<n>(s, 0, z), or all uniforms from draw_uniforms<n>(s, 0, u) (mixed branches: normals first, the
// uniforms from the Philox blocks after them)
static void emit_samples(std::ostringstream& o, const std::vector<SampleStmt>& st, int D) {
    int nn = 0, nu = 0;
    for (auto& s : st) { nn += s.dist == "normal"; nu += s.dist == "uniform"; }
    if (nn) o << "            Real z[" << nn << "];\n            draw_normals<" << nn << ">(s, 0, z);\n";
    if (nu) o << "            Real u[" << nu << "];\n            draw_uniforms<" << nu << ">(s, " << (nn ? "(uint32_t)((" + std::to_string(nn) + " + kPerBlock - 1) / kPerBlock)" : "0") << ", u);\n";
    int in = 0, iu = 0;
    for (int d = 0; d < D; ++d) {
        const SampleStmt& s = st[d];
        std::string v;
        if (s.dist == "normal") {   // normal.rs:26: z * std + mean
            v = is_zero_literal(s.a0) && s.add_to.empty() ? "z[" + std::to_string(in) + "] * (" + s.a1 + ")"
                                                          : "(z[" + std::to_string(in) + "] * (" + s.a1 + ") + (Real)(" + s.a0 + "))";
            ++in;
        } else if (s.dist == "uniform") {   // uniform.rs:28-32: u * (b - a) + a
            v = "u[" + std::to_string(iu) + "] * ((Real)(" + s.a1 + ") - (Real)(" + s.a0 + ")) + (Real)(" + s.a0 + ")";
            ++iu;
        } else v = "(Real)(" + s.a0 + ")";
        if (!s.add_to.empty()) v = "(" + s.add_to + ") + " + v;
        o << "            xn[" << d << "] = " << v << ";\n";
    }
}

static int generate(const Json& root, JitSpec& sp, std::string& err) {
    const Json* nm = root.get("name");
    sp.name = nm && nm->kind == Json::STR ? nm->str : "model";
    const Json *sd = root.get("state_dim"), *od = root.get("obs_dim");
    if (!sd || sd->kind != Json::NUM || !od || od->kind != Json::NUM) { err = "\"state_dim\" and \"obs_dim\" are required"; return -1; }
    sp.D = (int)sd->num; sp.nobs = (int)od->num;
    if (sp.D < 1 || sp.D > 8 || sp.nobs < 0 || sp.nobs > 4) { err = "state_dim must be 1..8 and obs_dim 0..4"; return -1; }
    if (const Json* ps = root.get("params")) {
        if (ps->kind != Json::OBJ) { err = "\"params\": an object of name: number"; return -1; }
        for (auto& kv : ps->obj) {
            if (kv.second.kind != Json::NUM) { err = "parameter '" + kv.first + "' must be a number"; return -1; }
            for (char c : kv.first) if (!(std::isalnum((unsigned char)c) || c == '_')) { err = "parameter names are identifiers"; return -1; }
            if (kv.first.empty() || std::isdigit((unsigned char)kv.first[0]) || kv.first == "x" || kv.first == "y" || kv.first == "t" || kv.first == "z" || kv.first == "u" || kv.first == "s")
                { err = "parameter name '" + kv.first + "' is reserved"; return -1; }
            sp.params.emplace_back(kv.first, kv.second.num);
        }
    }
    std::vector<SampleStmt> init, step;
    if (parse_samples(root.get("init"), sp.D, "init", init, err) || parse_samples(root.get("step"), sp.D, "step", step, err)) return -1;
    std::map<std::string, double> pval;
    for (auto& kv : sp.params) pval[kv.first] = kv.second;
    auto const_value = [&](const std::string& e, double& v) {   // a parameter name or a numeric literal: can be hoisted to the host
        auto it = pval.find(e);
        if (it != pval.end()) { v = it->second; return true; }
        char* end = nullptr;
        v = std::strtod(e.c_str(), &end);
        return end && end != e.c_str() && *end == 0;
    };
    for (int dt = 0; dt < 2; ++dt) {
        const bool f32 = dt == MPL_F32;
        auto real = [&](double v) { return f32 ? (double)(float)v : v; };
        std::vector<double>& der = sp.derived[dt];
        std::ostringstream o;
        o << "#include \"pf_kernels.cuh\"\nnamespace mpl {\ntemplate <typename Real> struct JitModel {\n";
        o << "    static constexpr int D = " << sp.D << ";\n    static constexpr int NOBS = " << sp.nobs << ";\n    static constexpr bool kGroupDraws = false;\n";
        o << "    static constexpr int kPerBlock = sizeof(Real) == 4 ? 4 : 2;   // normals / uniforms per Philox block\n";
        // observation terms: collect first (they decide how many derived constants there are)
        std::ostringstream obs_code;
        std::vector<std::string> pooled;     // z_i expressions of the pooled normal terms
        double pooled_const = 0.;
        int n_pooled = 0;
        const Json* ob = root.get("observe");
        if (ob && ob->kind != Json::ARR) { err = "\"observe\": a list of observation terms"; return -1; }
        std::vector<std::string> other_terms;
        for (size_t k = 0; ob && k < ob->arr.size(); ++k) {
            const Json& t = ob->arr[k];
            const Json* d = t.get("dist");
            if (t.kind != Json::OBJ || !d || d->kind != Json::STR) { err = "observe: {\"dist\": ...} expected"; return -1; }
            const std::string dist = d->str;
            auto strs = [&](const char* key, size_t n, std::vector<std::string>& out) {
                const Json* a = t.get(key);
                if (!a) return false;
                if (a->kind != Json::ARR) { std::string e; if (n != 1 || !expr_of(*a, e)) return false; out.push_back(e); return safe_expr(e); }
                if (a->arr.size() != n) return false;
                for (auto& x : a->arr) { std::string e; if (!expr_of(x, e) || !safe_expr(e)) return false; out.push_back(e); }
                return true;
            };
            std::vector<std::string> val, args;
            if (dist == "normal" || dist == "uniform") {
                if (!strs("value", 1, val) || !strs("args", 2, args)) { err = "observe " + dist + ": {\"value\": expr, \"args\": [a, b]}"; return -1; }
                double sdv;
                if (dist == "normal" && const_value(args[1], sdv)) {
                    // pooled like the built-in functors: z = (value - mean) * (1 / std) with 1 / std and the constant hoisted
                    const double sd_r = real(sdv);
                    der.push_back(f32 ? (double)(1.0f / (float)sd_r) : 1.0 / sd_r);
                    pooled.push_back("((Real)(" + val[0] + ") - (" + args[0] + ")) * (Real)c[" + std::to_string(der.size() - 1) + "]");
                    pooled_const += std::log(sd_r);
                    ++n_pooled;
                } else if (dist == "normal") other_terms.push_back("normal_logpdf<Real>((Real)(" + val[0] + "), (Real)(" + args[0] + "), (Real)(" + args[1] + "))");
                else other_terms.push_back("(Real)uniform_logpdf((double)(" + val[0] + "), (double)(" + args[0] + "), (double)(" + args[1] + "))");
            } else if (dist == "bernoulli") {
                if (!strs("value", 1, val) || !strs("args", 1, args)) { err = "observe bernoulli: {\"value\": expr, \"args\": [p]}"; return -1; }
                other_terms.push_back("(Real)bernoulli_logpdf((" + val[0] + ") != 0, (double)(" + args[0] + "))");
            } else if (dist == "mvnormal2") {
                std::vector<std::string> mean, cov;
                if (!strs("value", 2, val) || !strs("mean", 2, mean) || !strs("cov", 4, cov)) { err = "observe mvnormal2: {\"value\": [2], \"mean\": [2], \"cov\": [4]}"; return -1; }
                double cv[4];
                for (int i = 0; i < 4; ++i) if (!const_value(cov[i], cv[i])) { err = "mvnormal2: the covariance entries must be parameters or numbers (the determinant and inverse are hoisted)"; return -1; }
                const double det = cv[0] * cv[3] - cv[2] * cv[1];   // nalgebra 2x2 closed forms (mvnormal.rs:17-18), hoisted (quirk Q8)
                if (det == 0.) { err = "mvnormal2: singular covariance"; return -1; }
                const size_t b = der.size();
                der.push_back(cv[3] / det); der.push_back(-cv[1] / det); der.push_back(-cv[2] / det); der.push_back(cv[0] / det);
                der.push_back(2. * std::log(2. * kPi) + std::log(det));
                other_terms.push_back("(Real)mvnormal2_logpdf((double)(" + val[0] + "), (double)(" + val[1] + "), (double)(" + mean[0] + "), (double)(" + mean[1] + "), c + " +
                                      std::to_string(b) + ", c[" + std::to_string(b + 4) + "])");
            } else if (dist == "expr") {
                if (!strs("value", 1, val)) { err = "observe expr: {\"value\": log-density expression}"; return -1; }
                other_terms.push_back("(Real)(" + val[0] + ")");
            } else { err = "observe: unknown dist '" + dist + "'"; return -1; }
        }
        size_t pooled_const_idx = 0;
        if (n_pooled) { der.push_back(real((double)n_pooled * 0.5 * 1.8378770664093453 + pooled_const)); pooled_const_idx = der.size() - 1; }
        const size_t np = sp.params.size(), nc = der.size();
        o << "    double p[" << (np ? np : 1) << "];   // parameters\n    double c[" << (nc ? nc : 1) << "];   // derived constants (evaluated on the host in the kernel's precision)\n";
        o << "    __device__ __forceinline__ Real kernel(int64_t t, const Stream& s, Real (&x)[D], const Obs& obs) const {\n";
        for (size_t i = 0; i < np; ++i) o << "        const Real " << sp.params[i].first << " = (Real)p[" << i << "]; (void)" << sp.params[i].first << ";\n";
        o << "        const Real pi = (Real)3.14159265358979323846; (void)pi;\n";
        o << "        Real xn[D];\n        if (t == 0) {\n";
        emit_samples(o, init, sp.D);
        o << "        } else {\n";
        emit_samples(o, step, sp.D);
        o << "        }\n#pragma unroll\n        for (int d = 0; d < D; ++d) x[d] = xn[d];\n";
        o << "        const Real y[4] = {(Real)obs.v[0], (Real)obs.v[1], (Real)obs.v[2], (Real)obs.v[3]}; (void)y;\n";
        o << "        Real w = 0;\n";
        if (n_pooled) {
            for (int i = 0; i < n_pooled; ++i) o << "        const Real z" << i << "_ = " << pooled[i] << ";\n";
            o << "        w = (Real)-0.5 * (";
            for (int i = 0; i < n_pooled; ++i) o << (i ? " + " : "") << "z" << i << "_ * z" << i << "_";
            o << ") - (Real)c[" << pooled_const_idx << "];\n";
        }
        for (auto& t : other_terms) o << "        w += " << t << ";\n";
        o << "        return w;\n    }\n};\n}  // namespace mpl\n";
        sp.source[dt] = o.str();
    }
    return 0;
}

// ---- NVRTC (loaded on demand: the library itself does not link against it) --------------------------------------------------
struct Nvrtc {
    void* lib = nullptr;
    nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    nvrtcResult (*DestroyProgram)(nvrtcProgram*) = nullptr;
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*AddNameExpression)(nvrtcProgram, const char*) = nullptr;
    nvrtcResult (*GetLoweredName)(nvrtcProgram, const char*, const char**) = nullptr;
    const char* (*GetErrorString)(nvrtcResult) = nullptr;
    bool load(std::string& err) {
        if (lib) return true;
        for (const char* n : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"}) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
        if (!lib) { err = "libnvrtc.so.12 not found (the model front-end compiles specs with NVRTC)"; return false; }
#define MPL_SYM(f) *(void**)(&f) = dlsym(lib, "nvrtc" #f); if (!f) { err = "nvrtc" #f " missing"; return false; }
        MPL_SYM(CreateProgram) MPL_SYM(DestroyProgram) MPL_SYM(CompileProgram) MPL_SYM(GetProgramLogSize) MPL_SYM(GetProgramLog) MPL_SYM(GetCUBINSize) MPL_SYM(GetCUBIN)
        MPL_SYM(AddNameExpression) MPL_SYM(GetLoweredName) MPL_SYM(GetErrorString)
#undef MPL_SYM
        return true;
    }
};
static Nvrtc g_nvrtc;
static std::mutex g_jit_mutex;

// (mode, sharded, nested) combinations launch_extend uses
struct Variant { int mode; bool sharded; int nested; };
static const Variant kVariantsF32[] = {{EXT_INIT, false, 0}, {EXT_ACCUM, false, 0}, {EXT_GATHER, false, 0}, {EXT_GATHER, true, 0}, {EXT_DYNAMIC, false, 0}, {EXT_DYNAMIC, true, 0},
                                       {EXT_INIT, false, 1}, {EXT_GATHER, false, 1}, {EXT_GATHER, true, 1}, {EXT_INIT, false, 2}, {EXT_DYNAMIC, false, 2}, {EXT_DYNAMIC, true, 2}};
static const Variant kVariantsF64[] = {{EXT_INIT, false, 0}, {EXT_ACCUM, false, 0}, {EXT_GATHER, false, 0}, {EXT_GATHER, true, 0}, {EXT_DYNAMIC, false, 0}, {EXT_DYNAMIC, true, 0}};
static std::string variant_expr(int dtype, const Variant& v) {
    const char* real = dtype == MPL_F32 ? "float" : "double";
    std::ostringstream o;
    o << "mpl::pf_extend_kernel<mpl::JitModel<" << real << ">, " << real << ", " << v.mode << ", " << (v.sharded ? "true" : "false") << ", " << v.nested << ">";
    return o.str();
}

struct JitProgram {
    JitSpec spec;
    std::vector<char> cubin[2];
    std::map<std::string, std::string> lowered[2];   // name expression -> mangled kernel name
    std::string log[2];
    bool compiled[2] = {false, false};
    cudaLibrary_t library[2] = {nullptr, nullptr};   // loaded on first launch (needs a device)
    std::map<std::string, cudaKernel_t> kernels[2];
    ~JitProgram() { for (int d = 0; d < 2; ++d) if (library[d]) cudaLibraryUnload(library[d]); }
};

static int jit_compile(JitProgram& jp, int dtype) {
    std::lock_guard<std::mutex> lock(g_jit_mutex);
    if (jp.compiled[dtype]) return MPL_OK;
    std::string err;
    if (!g_nvrtc.load(err)) return fail(MPL_ERR_UNSUPPORTED, err);
    std::vector<const char*> hdr_text, hdr_name;
    for (const EmbeddedHeader& h : kHeaders) { hdr_name.push_back(h.name); hdr_text.push_back(h.text); }
    nvrtcProgram prog;
    nvrtcResult r = g_nvrtc.CreateProgram(&prog, jp.spec.source[dtype].c_str(), (jp.spec.name + ".cu").c_str(), (int)hdr_name.size(), hdr_text.data(), hdr_name.data());
    if (r != NVRTC_SUCCESS) return fail(MPL_ERR_CUDA, std::string("nvrtcCreateProgram: ") + g_nvrtc.GetErrorString(r));
    const size_t nv = dtype == MPL_F32 ? sizeof kVariantsF32 / sizeof(Variant) : sizeof kVariantsF64 / sizeof(Variant);
    const Variant* vs = dtype == MPL_F32 ? kVariantsF32 : kVariantsF64;
    std::vector<std::string> exprs;
    for (size_t i = 0; i < nv; ++i) { exprs.push_back(variant_expr(dtype, vs[i])); g_nvrtc.AddNameExpression(prog, exprs.back().c_str()); }
    const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "--device-int128", "-lineinfo"};
    r = g_nvrtc.CompileProgram(prog, 4, opts);
    size_t ls = 0;
    g_nvrtc.GetProgramLogSize(prog, &ls);
    jp.log[dtype].assign(ls ? ls - 1 : 0, 0);
    if (ls > 1) g_nvrtc.GetProgramLog(prog, &jp.log[dtype][0]);
    if (r != NVRTC_SUCCESS) {
        std::string msg = std::string("model spec '") + jp.spec.name + "' does not compile (" + g_nvrtc.GetErrorString(r) + "):\n" + jp.log[dtype];
        g_nvrtc.DestroyProgram(&prog);
        return fail(MPL_ERR_INVALID, msg);
    }
    size_t cs = 0;
    g_nvrtc.GetCUBINSize(prog, &cs);
    jp.cubin[dtype].resize(cs);
    g_nvrtc.GetCUBIN(prog, jp.cubin[dtype].data());
    for (auto& e : exprs) {
        const char* low = nullptr;
        if (g_nvrtc.GetLoweredName(prog, e.c_str(), &low) != NVRTC_SUCCESS || !low) { g_nvrtc.DestroyProgram(&prog); return fail(MPL_ERR_CUDA, "nvrtcGetLoweredName failed for " + e); }
        jp.lowered[dtype][e] = low;
    }
    g_nvrtc.DestroyProgram(&prog);
    jp.compiled[dtype] = true;
    return MPL_OK;
}

// the kernel of one (mode, sharded, nested) variant, loading the module on first use
static int jit_kernel(JitProgram& jp, int dtype, const Variant& v, cudaKernel_t* out) {
    int rc = jit_compile(jp, dtype);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(g_jit_mutex);
    if (!jp.library[dtype]) MPL_CUDA_OK(cudaLibraryLoadData(&jp.library[dtype], jp.cubin[dtype].data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
    const std::string e = variant_expr(dtype, v);
    auto it = jp.kernels[dtype].find(e);
    if (it == jp.kernels[dtype].end()) {
        auto low = jp.lowered[dtype].find(e);
        if (low == jp.lowered[dtype].end()) return fail(MPL_ERR_UNSUPPORTED, "this kernel variant is not compiled for run-time models: " + e);
        cudaKernel_t k;
        MPL_CUDA_OK(cudaLibraryGetKernel(&k, jp.library[dtype], low->second.c_str()));
        it = jp.kernels[dtype].emplace(e, k).first;
    }
    *out = it->second;
    return MPL_OK;
}

// called by pf.cu's extend dispatch for models created by mpl_model_compile: same ExtendArgs, same launch attributes
int jit_launch_extend(const mpl_model& m, int dtype, int mode, bool sharded, int nested, const void* extend_args, unsigned int grid, unsigned int block, cudaStream_t stream, bool pdl) {
    if (!m.jit) return fail(MPL_ERR_INVALID, "not a run-time model");
    JitProgram& jp = *m.jit;
    cudaKernel_t k;
    int rc = jit_kernel(jp, dtype, Variant{mode, sharded, nested}, &k);
    if (rc) return rc;
    // the functor, as the generated struct lays it out: double p[max(np, 1)], double c[max(nc, 1)]
    std::vector<double> functor;
    for (auto& kv : jp.spec.params) functor.push_back(kv.second);
    if (jp.spec.params.empty()) functor.push_back(0.);
    for (double d : jp.spec.derived[dtype]) functor.push_back(d);
    if (jp.spec.derived[dtype].empty()) functor.push_back(0.);
    void* args[2] = {const_cast<void*>(extend_args), functor.data()};
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    MPL_CUDA_OK(cudaLaunchKernelExC(&cfg, (const void*)k, args));
    return MPL_OK;
}

}  // namespace mpl

using namespace mpl;

extern "C" mpl_model* mpl_model_compile(const char* spec_json) {
    if (!spec_json) { fail(MPL_ERR_INVALID, "null spec"); return nullptr; }
    JsonReader rd{spec_json, spec_json + std::strlen(spec_json), {}};
    Json root;
    if (!rd.value(root) || root.kind != Json::OBJ) { fail(MPL_ERR_INVALID, "model spec: " + (rd.err.empty() ? std::string("a JSON object is expected") : rd.err)); return nullptr; }
    auto jp = std::make_shared<JitProgram>();
    std::string err;
    if (generate(root, jp->spec, err)) { fail(MPL_ERR_INVALID, "model spec: " + err); return nullptr; }
    auto* m = new mpl_model;
    m->kind = M_JIT; m->name = jp->spec.name;
    for (auto& kv : jp->spec.params) m->params.push_back(kv.second);
    m->state_dim = jp->spec.D; m->obs_dim = jp->spec.nobs; m->num_latents = 0;
    m->jit = jp;
    return m;
}

// compile now (otherwise the first step does): log receives the compiler's messages.  Needs no device.
extern "C" int mpl_model_jit_compile(mpl_model* m, int dtype, char* log, size_t log_bytes) {
    if (!m || !m->jit || (dtype != MPL_F32 && dtype != MPL_F64)) return fail(MPL_ERR_INVALID, "a model created by mpl_model_compile and MPL_F32 / MPL_F64 are expected");
    const int rc = jit_compile(*m->jit, dtype);
    if (log && log_bytes) { std::strncpy(log, m->jit->log[dtype].c_str(), log_bytes - 1); log[log_bytes - 1] = 0; }
    return rc;
}

// the generated translation unit (what NVRTC compiles), for inspection
extern "C" const char* mpl_model_jit_source(const mpl_model* m, int dtype) {
    if (!m || !m->jit || (dtype != MPL_F32 && dtype != MPL_F64)) return nullptr;
    return m->jit->spec.source[dtype].c_str();
}
