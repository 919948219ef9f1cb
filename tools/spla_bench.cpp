// tools/spla_bench.cpp -- the bench workload through spla's OWN public C++ API (spla.hpp), on the CUDA backend of this repository
// and on spla's CPU backend in the same process: what a user of the reference sees after switching backends, set-up included.
//
//   spla_bench <prefix> <n> <nnz> [steps=20] [cpu_steps=2] [pr=1]
//
// <prefix>_Ai.bin / _Aj.bin / _Ax.bin hold the row-sorted COO of the matrix (uint32, uint32, float32: the arrays Matrix::build
// takes, reference include/spla/matrix.hpp), written by bench.py from the SAME graph its GPU arm times. Reported as one JSON line:
//   build_ms            Matrix::build (host: COO -> the CPU decoration)
//   first_call_ms       the first exec_mxv_masked on the CUDA backend: format conversions of the storage manager
//                       (validate_rw: CpuCoo -> ... -> AccCsr, reference src/storage/storage_manager.hpp:120-222), the H2D copies
//                       and the handle build (column classes), everything a steady-state call no longer pays
//   step_ms             steady state: exec_mxv_masked(r, mask, A, v, MULT, PLUS, NQZERO, 0) K times back to back, closed by one
//                       exec_v_count_mf (a 4-byte read: the only way to wait for the device through the API); count_ms subtracted
//   cpu_step_ms         the same call on spla's CPU backend (Library::set_force_no_acceleration(true)), same matrix object
//   pr_*                spla::pr(alpha 0.85, eps 1e-6) on both backends: time, max relative difference of the ranks
// Built by spla_b200/src/Makefile into spla_b200/lib/spla_bench (needs the reference checkout at build time only).
#include <spla.hpp>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace spla;
using Clock = std::chrono::steady_clock;
static double ms_since(Clock::time_point t0) { return std::chrono::duration<double, std::milli>(Clock::now() - t0).count(); }

template<typename T>
static bool read_file(const std::string& path, std::vector<T>& out, std::size_t count) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    out.resize(count);
    const std::size_t got = std::fread(out.data(), sizeof(T), count, f);
    std::fclose(f);
    return got == count;
}

int main(int argc, char** argv) {
    setvbuf(stdout, nullptr, _IOLBF, 0);
    if (argc < 4) {
        std::printf("usage: spla_bench <prefix> <n> <nnz> [steps] [cpu_steps] [pr]\n");
        return 2;
    }
    const std::string prefix = argv[1];
    const uint        n      = uint(std::strtoul(argv[2], nullptr, 10));
    const std::size_t nnz    = std::strtoull(argv[3], nullptr, 10);
    const int         steps  = argc > 4 ? std::atoi(argv[4]) : 20;
    const int         csteps = argc > 5 ? std::atoi(argv[5]) : 2;
    const bool        do_pr  = argc > 6 ? std::atoi(argv[6]) != 0 : true;

    Library*    lib = Library::get();
    std::string info;
    lib->get_accelerator_info(info);
    if (info.find("CUDA") == std::string::npos) {
        std::printf("{\"error\": \"the CUDA accelerator is not active: %s\"}\n", info.c_str());
        return 2;
    }
    std::vector<uint>  Ai, Aj;
    std::vector<float> Ax;
    if (!read_file(prefix + "_Ai.bin", Ai, nnz) || !read_file(prefix + "_Aj.bin", Aj, nnz) || !read_file(prefix + "_Ax.bin", Ax, nnz)) {
        std::printf("{\"error\": \"cannot read %s_*.bin\"}\n", prefix.c_str());
        return 2;
    }
    auto t0 = Clock::now();
    auto A  = Matrix::make(n, n, FLOAT);
    A->build(MemView::make(Ai.data(), nnz * sizeof(uint)), MemView::make(Aj.data(), nnz * sizeof(uint)), MemView::make(Ax.data(), nnz * sizeof(float)));
    const double build_ms = ms_since(t0);
    std::vector<uint>().swap(Ai);
    std::vector<uint>().swap(Aj);
    std::vector<float>().swap(Ax);

    auto zero = Scalar::make_float(0.0f);
    auto cnt  = Scalar::make_uint(0);
    auto run  = [&](bool cpu, int k, double& first_ms, double& step_ms, double& count_ms, double& checksum) {
        lib->set_force_no_acceleration(cpu);
        auto v = Vector::make(n, FLOAT), mask = Vector::make(n, FLOAT), r = Vector::make(n, FLOAT);
        v->fill_with(Scalar::make_float(1.0f / float(n)));
        mask->fill_with(Scalar::make_float(1.0f));
        auto t = Clock::now();
        exec_mxv_masked(r, mask, A, v, MULT_FLOAT, PLUS_FLOAT, NQZERO_FLOAT, zero);
        exec_v_count_mf(cnt, r);
        first_ms = ms_since(t);
        t        = Clock::now();
        exec_v_count_mf(cnt, r);
        count_ms = ms_since(t);
        t        = Clock::now();
        for (int i = 0; i < k; ++i) exec_mxv_masked(r, mask, A, v, MULT_FLOAT, PLUS_FLOAT, NQZERO_FLOAT, zero);
        exec_v_count_mf(cnt, r);
        step_ms = (ms_since(t) - count_ms) / double(k);
        auto sum = Scalar::make_float(0.0f);
        exec_v_reduce(sum, zero, r, PLUS_FLOAT);
        checksum = double(sum->as_float());
    };
    double cu_first = 0, cu_step = 0, cu_count = 0, cu_sum = 0, cp_first = 0, cp_step = 0, cp_count = 0, cp_sum = 0;
    run(false, steps, cu_first, cu_step, cu_count, cu_sum);
    if (csteps > 0) run(true, csteps, cp_first, cp_step, cp_count, cp_sum);

    double pr_cuda_ms = 0, pr_cpu_ms = 0, pr_maxrel = 0;
    if (do_pr) {
        std::vector<float> res[2];
        for (int pass = 0; pass < (csteps > 0 ? 2 : 1); ++pass) {
            lib->set_force_no_acceleration(pass == 1);
            auto p = Vector::make(n, FLOAT);
            auto t = Clock::now();
            pr(p, A, 0.85f, 1e-6f);
            exec_v_count_mf(cnt, p);
            (pass == 0 ? pr_cuda_ms : pr_cpu_ms) = ms_since(t);
            res[pass].resize(n);
            // one bulk read through the public API (keys + values views)
            ref_ptr<MemView> keys, vals;
            p->read(keys, vals);
            const std::size_t stored = vals->get_size() / sizeof(float);
            const uint*       k      = static_cast<const uint*>(keys->get_buffer());
            const float*      x      = static_cast<const float*>(vals->get_buffer());
            std::fill(res[pass].begin(), res[pass].end(), 0.0f);
            for (std::size_t q = 0; q < stored; ++q) res[pass][k[q]] = x[q];
        }
        if (csteps > 0)
            for (uint i = 0; i < n; ++i)
                if (res[1][i] != 0.f) pr_maxrel = std::max(pr_maxrel, double(std::fabs(res[0][i] - res[1][i]) / std::fabs(res[1][i])));
    }
    std::printf("{\"accelerator\": \"%s\", \"n\": %u, \"nnz\": %zu, \"build_ms\": %.2f, \"first_call_ms\": %.2f, \"step_ms\": %.4f, \"count_ms\": %.4f, "
                "\"gteps\": %.3f, \"sum_r\": %.9g, \"steps\": %d, \"cpu_first_call_ms\": %.2f, \"cpu_step_ms\": %.2f, \"cpu_gteps\": %.4f, \"cpu_sum_r\": %.9g, "
                "\"cpu_steps\": %d, \"pr_cuda_ms\": %.2f, \"pr_cpu_ms\": %.2f, \"pr_max_rel_diff\": %.3e}\n",
                info.c_str(), n, nnz, build_ms, cu_first, cu_step, cu_count, double(nnz) / cu_step / 1e6, cu_sum, steps, cp_first, cp_step,
                cp_step > 0 ? double(nnz) / cp_step / 1e6 : 0.0, cp_sum, csteps, pr_cuda_ms, pr_cpu_ms, pr_maxrel);
    return 0;
}
