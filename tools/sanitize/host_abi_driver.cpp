// sanitizer harness: the host half of the C ABI over a list of expressions (one per line on stdin)
#include <cstdio>
#include <iostream>
#include <string>
#include <vector>
#include "zignal_b200.h"
int main() {
    std::string line;
    long n = 0, ok = 0;
    while (std::getline(std::cin, line)) {
        ++n;
        int a = 0, b = 0, cnt = 0, tup = 0;
        int d[64];
        char buf[1 << 16];
        zg_expr_arity(line.c_str(), &a, &b);
        zg_expr_delays(line.c_str(), 0, d, 64, &cnt);
        zg_expr_delays(line.c_str(), 1, d, 64, &cnt);
        zg_expr_canonical(line.c_str(), buf, sizeof buf);
        int in_t[16] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1}, out_t[64];
        zg_expr_result_types(line.c_str(), in_t, a < 16 ? a : 16, out_t, 64, &cnt, &tup);
        zg_graph* g = nullptr;
        if (zg_graph_compile(line.c_str(), &g) != ZG_OK) continue;
        ++ok;
        zg_graph_info gi;
        zg_graph_get_info(g, &gi);
        zg_graph_kernel_class(g, buf, sizeof buf);
        int lin = 0;
        zg_graph_linearity(g, &lin);
        if (gi.n_state <= 64 && gi.n_params <= 8) {              // the linear-tick analysis behind time_parallel
            std::vector<double> A((size_t)gi.n_state * gi.n_state + 1);
            const float prm[8] = {0.5f, -0.25f, 0.9f, 1.0f, -1.5f, 0.f, 2.f, 0.125f};
            int K = 0;
            zg_graph_state_matrix(g, prm, gi.n_params, A.data(), A.size());
            zg_graph_settling_time(g, prm, gi.n_params, 128, 1024, 1e-9, &K);
        }
        (void)zg_graph_dump(g);
        zg_voice* v = nullptr;
        if (zg_voice_create(g, &v) == ZG_OK) {
            double in[8] = {1, -2, 3, 0.5, 2, 1, 1, 1}, out[16];
            int idt[8] = {0, 1, 2, 1, 0, 1, 2, 1}, odt[16];
            for (int t = 0; t < 4; ++t) zg_voice_tick(v, in, idt, out, odt);
            const double edge[8] = {-2147483648.0, -1, 3e10, -3e10, 0.0 / 0.0, 1e300, -1, 0};   // INT_MIN / -1, out-of-range ints, NaN
            const int all_int[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            zg_voice_tick(v, edge, all_int, out, odt);
            zg_voice* w = nullptr;
            if (zg_voice_clone(v, &w) == ZG_OK) { zg_voice_tick(w, in, idt, out, odt); zg_voice_destroy(w); }
            zg_voice_destroy(v);
        }
        zg_graph_destroy(g);
    }
    std::printf("%ld expressions, %ld valid graphs\n", n, ok);
    return 0;
}
