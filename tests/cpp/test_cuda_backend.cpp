// tests/cpp/test_cuda_backend.cpp -- differential test of spla's CUDA backend against spla's own CPU backend, in ONE process,
// through the unchanged public C++ API (spla.hpp): the same exec_mxv_masked / exec_vxm_masked / bfs / sssp / pr calls run once
// with Library::set_force_no_acceleration(true) (CPU algorithms, the oracle) and once on the CUDA accelerator, the pattern of the
// reference's examples/bfs.cpp:85-107. Integer / MIN-MAX / logical results must be bit-identical, float sums within 1e-5
// relative. Prints one line per check and exits non-zero on the first mismatch.
#include <spla.hpp>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <set>
#include <string>
#include <vector>

using namespace spla;

static int g_failed = 0;
static int g_checks = 0;

static void check(bool ok, const std::string& what) {
    ++g_checks;
    if (!ok) ++g_failed;
    std::printf("[%s] %s\n", ok ? " OK " : "FAIL", what.c_str());
    std::fflush(stdout);
}

struct Graph {
    uint                            n;
    std::vector<uint>               Ai, Aj;// row-sorted, symmetric, no loops, no duplicates
    std::vector<std::vector<uint>>  adj;
};

// small R-MAT-like generator (a,b,c,d) = (0.57,0.19,0.19,0.05), symmetrised + dedup + no loops like MtxLoader (reference src/io.cpp:159-214)
static Graph make_rmat(int scale, int edge_factor, unsigned seed) {
    Graph g;
    g.n = 1u << scale;
    std::mt19937                          rng(seed);
    std::uniform_real_distribution<float> uni(0.f, 1.f);
    std::set<std::pair<uint, uint>>       edges;
    for (std::size_t e = 0; e < std::size_t(edge_factor) << scale; ++e) {
        uint i = 0, j = 0;
        for (int b = 0; b < scale; ++b) {
            const float r = uni(rng);
            i = (i << 1) | (r >= 0.76f);
            j = (j << 1) | ((r >= 0.57f && r < 0.76f) || r >= 0.95f);
        }
        if (i == j) continue;
        edges.insert({i, j});
        edges.insert({j, i});
    }
    g.adj.resize(g.n);
    for (auto& e : edges) {
        g.Ai.push_back(e.first);
        g.Aj.push_back(e.second);
        g.adj[e.first].push_back(e.second);
    }
    return g;
}

template<typename T>
static ref_ptr<Matrix> build_matrix(const Graph& g, const ref_ptr<Type>& type, const std::vector<T>& values) {
    auto A = Matrix::make(g.n, g.n, type);
    A->build(MemView::make((void*) g.Ai.data(), g.Ai.size() * sizeof(uint)), MemView::make((void*) g.Aj.data(), g.Aj.size() * sizeof(uint)),
             MemView::make((void*) values.data(), values.size() * sizeof(T)));
    return A;
}

static std::vector<int> read_int(const ref_ptr<Vector>& v) {
    std::vector<int> out(v->get_n_rows());
    for (uint i = 0; i < out.size(); ++i) v->get_int(i, out[i]);
    return out;
}
static std::vector<float> read_float(const ref_ptr<Vector>& v) {
    std::vector<float> out(v->get_n_rows());
    for (uint i = 0; i < out.size(); ++i) v->get_float(i, out[i]);
    return out;
}
// north_star: |cuda - cpu| <= rtol * |cpu| per element, no absolute floor. An element beyond that passes only if the CUDA value is no
// further from the float64 value `exact` than the reference's own fp32 result is (or within rtol of exact): the reference's
// left-to-right fp32 fold of a hub row loses up to d * 2^-24 relative, the device's segmented tree does not. Both maxima are printed.
static bool close_rel(const std::vector<float>& a, const std::vector<float>& b, float rtol, const std::vector<double>* exact = nullptr) {
    double      worst = 0.0, worst_dev = 0.0, worst_ref = 0.0;
    bool        ok = true;
    std::size_t escaped = 0;
    for (std::size_t i = 0; i < a.size(); ++i) {
        if (a[i] == b[i]) continue;
        const double d   = std::fabs(double(a[i]) - double(b[i]));
        const double rel = b[i] != 0.f ? d / std::fabs(double(b[i])) : INFINITY;
        worst            = std::max(worst, rel);
        if (d <= double(rtol) * std::fabs(double(b[i]))) continue;
        bool pass = false;
        if (exact) {
            const double x = (*exact)[i], de = std::fabs(double(a[i]) - x), re = std::fabs(double(b[i]) - x);
            pass      = de <= re || de <= double(rtol) * std::fabs(x);
            worst_dev = std::max(worst_dev, de / std::fabs(x));
            worst_ref = std::max(worst_ref, re / std::fabs(x));
            escaped += pass;
        }
        if (!pass) {
            if (ok) std::printf("   mismatch at %zu: %.9g vs %.9g\n", i, a[i], b[i]);
            ok = false;
        }
    }
    std::printf("   max relative difference cuda vs cpu %.3e (bar %.1e)", worst, double(rtol));
    if (escaped) std::printf("; %zu element(s) beyond the bar accepted: vs float64 the cuda result is off by <= %.3e, the cpu result by up to %.3e", escaped, worst_dev, worst_ref);
    std::printf("\n");
    return ok;
}

static void use_cpu(bool cpu) { Library::get()->set_force_no_acceleration(cpu); }

int main(int argc, char** argv) {
    setvbuf(stdout, nullptr, _IOLBF, 0);
    const int scale = argc > 1 ? std::atoi(argv[1]) : 12;
    Library*  lib   = Library::get();
    std::string info;
    lib->get_accelerator_info(info);
    std::printf("accelerator: %s\n", info.c_str());
    if (info.find("CUDA") == std::string::npos && info.find("cuda") == std::string::npos && !std::getenv("SPLA_TEST_DRY_RUN")) {
        std::printf("[FAIL] the CUDA accelerator is not active: this test has no CPU fallback\n");
        return 2;
    }

    Graph g = make_rmat(scale, 8, 42);
    std::printf("graph: n=%u nnz=%zu\n", g.n, g.Ai.size());
    const uint N = g.n;
    uint       source = 0;
    for (uint i = 0; i < N; ++i)
        if (g.adj[i].size() > g.adj[source].size()) source = i;

    // ---- mxv / vxm with the semirings of bfs / sssp / pr and a few argument-order sensitive ones ----------------------
    {
        std::mt19937 rng(7);
        std::vector<int> ones(g.Ai.size());
        for (auto& x : ones) x = 1 + int(rng() % 3);
        auto A = build_matrix<int>(g, INT, ones);
        struct Case { ref_ptr<OpBinary> m, a; ref_ptr<OpSelect> s; bool ee; const char* name; };
        Case cases[] = {{MULT_INT, PLUS_INT, EQZERO_INT, false, "INT MULT/PLUS/EQZERO"},
                        {BAND_INT, BOR_INT, EQZERO_INT, true, "INT BAND/BOR/EQZERO early-exit"},
                        {LAND_INT, LOR_INT, GTZERO_INT, false, "INT LAND/LOR/GTZERO"},
                        {PLUS_INT, MIN_INT, ALWAYS_INT, false, "INT PLUS/MIN/ALWAYS"},
                        {MINUS_INT, PLUS_INT, NQZERO_INT, false, "INT MINUS/PLUS/NQZERO"},
                        {FIRST_INT, SECOND_INT, ALWAYS_INT, false, "INT FIRST/SECOND/ALWAYS (ordered fold)"},
                        {MULT_INT, MAX_INT, LEZERO_INT, true, "INT MULT/MAX/LEZERO early-exit"}};
        for (auto& c : cases) {
            std::vector<int> res[2];
            std::vector<int> resx[2];
            for (int pass = 0; pass < 2; ++pass) {
                use_cpu(pass == 0);
                auto v = Vector::make(N, INT), mask = Vector::make(N, INT), r = Vector::make(N, INT), f = Vector::make(N, INT), rx = Vector::make(N, INT);
                std::mt19937 vr(99);
                for (uint i = 0; i < N; ++i) {
                    if (vr() % 3 == 0) v->set_int(i, int(vr() % 5) - 1);
                    if (vr() % 2 == 0) mask->set_int(i, int(vr() % 3) - 1);
                    if (vr() % 17 == 0) f->set_int(i, int(vr() % 4));// explicit zeros stay in the frontier (SURVEY 8a note F)
                }
                auto desc = Descriptor::make();
                desc->set_early_exit(c.ee);
                exec_mxv_masked(r, mask, A, v, c.m, c.a, c.s, Scalar::make_int(c.ee ? 0 : 2), desc);
                exec_vxm_masked(rx, mask, f, A, c.m, c.a, c.s, Scalar::make_int(0), desc);
                res[pass]  = read_int(r);
                resx[pass] = read_int(rx);
            }
            check(res[0] == res[1], std::string("mxv_masked ") + c.name + " : cuda == cpu (bit-exact)");
            check(resx[0] == resx[1], std::string("vxm_masked ") + c.name + " : cuda == cpu (bit-exact)");
        }
    }

    // ---- bfs: push, pull, push-pull (reference src/algorithm.cpp:45-120) ---------------------------------------------------
    {
        std::vector<int> ones(g.Ai.size(), 1);
        auto             A = build_matrix<int>(g, INT, ones);
        std::vector<int> ref(N);
        bfs_naive(ref, g.adj, source, ref_ptr<Descriptor>());
        for (int mode = 0; mode < 3; ++mode) {
            auto desc = Descriptor::make();
            desc->set_traversal_mode(static_cast<Descriptor::TraversalMode>(mode));
            desc->set_front_factor(0.05f);
            std::vector<int> res[2];
            for (int pass = 0; pass < 2; ++pass) {
                use_cpu(pass == 0);
                auto v = Vector::make(N, INT);
                bfs(v, A, source, desc);
                res[pass] = read_int(v);
            }
            const char* names[] = {"push", "pull", "push-pull"};
            check(res[0] == res[1], std::string("bfs ") + names[mode] + " : cuda depths == cpu depths (bit-exact)");
            check(res[1] == ref, std::string("bfs ") + names[mode] + " : cuda depths == bfs_naive");
        }
    }

    // ---- sssp (reference src/algorithm.cpp:158-229): MIN / PLUS on float, every candidate is one fp add => bit-exact ----
    {
        std::mt19937                          rng(3);
        std::uniform_real_distribution<float> w(1.f, 2.f);
        std::vector<float>                    weights(g.Ai.size());
        for (auto& x : weights) x = w(rng);
        auto A = build_matrix<float>(g, FLOAT, weights);
        for (int mode = 0; mode < 3; ++mode) {
            auto desc = Descriptor::make();
            desc->set_traversal_mode(static_cast<Descriptor::TraversalMode>(mode));
            desc->set_front_factor(0.05f);
            std::vector<float> res[2];
            for (int pass = 0; pass < 2; ++pass) {
                use_cpu(pass == 0);
                auto v = Vector::make(N, FLOAT);
                sssp(v, A, source, desc);
                res[pass] = read_float(v);
            }
            const char* names[] = {"push", "pull", "push-pull"};
            check(std::memcmp(res[0].data(), res[1].data(), N * sizeof(float)) == 0, std::string("sssp ") + names[mode] + " : cuda distances == cpu distances (bit-exact)");
        }
    }

    // ---- pagerank (reference src/algorithm.cpp:278-335): float PLUS => 1e-5 relative ---------------------------------------
    {
        std::vector<float> weights(g.Ai.size());
        for (std::size_t k = 0; k < g.Ai.size(); ++k) weights[k] = 0.85f / float(g.adj[g.Ai[k]].size());
        auto               A = build_matrix<float>(g, FLOAT, weights);
        std::vector<float> res[2];
        for (int pass = 0; pass < 2; ++pass) {
            use_cpu(pass == 0);
            auto p = Vector::make(N, FLOAT);
            pr(p, A, 0.85f, 1e-6f);
            res[pass] = read_float(p);
        }
        // the same loop in float64 on the fp32 matrix values (reference src/algorithm.cpp:278-335); the iteration count is the one
        // whose float64 iterate is closest to the cpu result
        std::vector<double> p64(N, double(1.0f / float(N))), nxt(N), best;
        const double        add = double((1.0f - 0.85f) / float(N));
        double              best_d = 1e300;
        for (int it = 0; it < 200; ++it) {
            std::fill(nxt.begin(), nxt.end(), add);
            for (std::size_t k = 0; k < g.Ai.size(); ++k) nxt[g.Ai[k]] += double(weights[k]) * p64[g.Aj[k]];
            p64.swap(nxt);
            double d = 0.0;
            for (uint i = 0; i < N; ++i) d = std::max(d, std::fabs(p64[i] - double(res[0][i])));
            if (d < best_d) best_d = d, best = p64;
            else if (d > 4 * best_d) break;
        }
        check(close_rel(res[1], res[0], 1e-5f, &best), "pr : cuda ranks within 1e-5 relative of cpu ranks per element (or at least as close to the float64 ranks as the cpu's)");
    }

    // ---- user-defined ops (reference src/op.cpp:294-342, tests/test_vector.cpp:285-315): compiled with NVRTC inside the backend, never
    //      handed to the cpu on the mxv / vxm path; the sequential semantics make them bit-exact against the cpu backend --------------------
    if (!std::getenv("SPLA_TEST_DRY_RUN")) {
        std::mt19937     rng(11);
        std::vector<int> vals(g.Ai.size());
        for (auto& x : vals) x = 1 + int(rng() % 3);
        auto A        = build_matrix<int>(g, INT, vals);
        auto my_plus  = OpBinary::make_int("my_plus1", "(int a, int b) { return a + b + 1; }", [](int a, int b) { return a + b + 1; });
        auto my_mult  = OpBinary::make_int("my_mult2", "(int a, int b) { return 2 * a * b - b; }", [](int a, int b) { return 2 * a * b - b; });
        auto my_sel   = OpSelect::make_int("my_odd", "(int a) { return (a & 1) != 0; }", [](int a) { return (a & 1) != 0; });
        struct UCase { ref_ptr<OpBinary> m, a; ref_ptr<OpSelect> s; bool ee; const char* name; };
        UCase ucases[] = {{MULT_INT, my_plus, EQZERO_INT, false, "user op_add"},
                          {my_mult, PLUS_INT, NQZERO_INT, false, "user op_mult"},
                          {my_mult, my_plus, my_sel, false, "user op_mult, op_add and op_select"},
                          {MULT_INT, my_plus, my_sel, true, "user op_add + op_select, early exit"}};
        for (auto& c : ucases) {
            std::vector<int> res[2], resx[2];
            Status           st_m[2], st_x[2];
            for (int pass = 0; pass < 2; ++pass) {
                use_cpu(pass == 0);
                auto v = Vector::make(N, INT), mask = Vector::make(N, INT), r = Vector::make(N, INT), f = Vector::make(N, INT), rx = Vector::make(N, INT);
                std::mt19937 vr(5);
                for (uint i = 0; i < N; ++i) {
                    if (vr() % 3 == 0) v->set_int(i, int(vr() % 5) - 1);
                    if (vr() % 2 == 0) mask->set_int(i, int(vr() % 3) - 1);
                    if (vr() % 13 == 0) f->set_int(i, int(vr() % 4));
                }
                auto desc = Descriptor::make();
                desc->set_early_exit(c.ee);
                st_m[pass] = exec_mxv_masked(r, mask, A, v, c.m, c.a, c.s, Scalar::make_int(c.ee ? 0 : 2), desc);
                st_x[pass] = exec_vxm_masked(rx, mask, f, A, c.m, c.a, c.s, Scalar::make_int(0), desc);
                res[pass]  = read_int(r);
                resx[pass] = read_int(rx);
            }
            check(st_m[1] == Status::Ok && st_x[1] == Status::Ok, std::string("mxv / vxm with ") + c.name + " : Status::Ok on the cuda backend (NVRTC)");
            check(res[0] == res[1], std::string("mxv_masked with ") + c.name + " : cuda == cpu (bit-exact)");
            check(resx[0] == resx[1], std::string("vxm_masked with ") + c.name + " : cuda == cpu (bit-exact)");
        }
        // float: mult and add must round separately (no fused multiply-add), as the two std::function calls of the cpu backend do
        {
            std::vector<float> w(g.Ai.size());
            std::mt19937       wr(3);
            for (auto& x : w) x = 0.5f + float(wr() % 1000) / 777.0f;
            auto Af      = build_matrix<float>(g, FLOAT, w);
            auto blend   = OpBinary::make_float("blend", "(float a, float b) { return 0.25f * a + 0.75f * b; }", [](float a, float b) { return 0.25f * a + 0.75f * b; });
            std::vector<float> res[2];
            for (int pass = 0; pass < 2; ++pass) {
                use_cpu(pass == 0);
                auto v = Vector::make(N, FLOAT), mask = Vector::make(N, FLOAT), r = Vector::make(N, FLOAT);
                std::mt19937 vr(17);
                for (uint i = 0; i < N; ++i)
                    if (vr() % 2 == 0) v->set_float(i, float(vr() % 100) / 31.0f);
                exec_mxv_masked(r, mask, Af, v, MULT_FLOAT, blend, EQZERO_FLOAT, Scalar::make_float(1.5f));
                res[pass] = read_float(r);
            }
            check(std::memcmp(res[0].data(), res[1].data(), N * sizeof(float)) == 0, "mxv_masked FLOAT with a user op_add (sequential fold, no fma) : cuda == cpu (bit-exact)");
        }
        // a source that does not compile is reported, not hidden behind the cpu
        {
            use_cpu(false);
            auto broken = OpBinary::make_int("broken", "(int a, int b) { return a +* b; }", [](int a, int b) { return a + b; });
            auto v = Vector::make(N, INT), mask = Vector::make(N, INT), r = Vector::make(N, INT);
            v->set_int(0, 1);
            Status st = exec_mxv_masked(r, mask, A, v, MULT_INT, broken, EQZERO_INT, Scalar::make_int(0));
            check(st == Status::CompilationError, "mxv_masked with a user op whose source does not compile returns CompilationError on the cuda backend");
        }
    }

    std::printf("%d checks, %d failed\n", g_checks, g_failed);
    lib->finalize();
    return g_failed ? 1 : 0;
}
