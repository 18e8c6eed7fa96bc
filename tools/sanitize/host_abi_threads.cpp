// ThreadSanitizer harness: one immutable graph shared by 8 threads (include/zignal_b200.h: "graph handles are immutable
// and shareable"), each with its own voices, ticking with different argument-type signatures (the per-signature tick
// programs are built lazily inside the shared graph), compiling graphs of its own and provoking errors (zg_last_error is
// thread local).
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
#include "zignal_b200.h"
int main() {
    zg_graph* g = nullptr;
    if (zg_graph_compile("(_1 + _2 , _1 - _2[_1]) |= (~(_2 + 0.5f*_1[_1]) | (_1 - 0.25*_1[_2]))", &g) != ZG_OK) return 1;
    std::vector<std::thread> th;
    std::vector<int> bad(8, 0);
    for (int t = 0; t < 8; ++t)
        th.emplace_back([&, t] {
            for (int rep = 0; rep < 200; ++rep) {
                zg_voice* v = nullptr;
                if (zg_voice_create(g, &v) != ZG_OK) { bad[t]++; continue; }
                double in[2] = {1.0 + t, 2.0}, out[4], first = 0;
                int idt[2] = {(t + rep) % 3, (t / 3 + rep) % 3}, odt[4];
                for (int k = 0; k < 8; ++k) {
                    if (zg_voice_tick(v, in, idt, out, odt) != ZG_OK) bad[t]++;
                    if (k == 0) first = out[0];
                }
                zg_voice* w = nullptr;
                zg_voice_clone(v, &w);
                zg_voice_destroy(w);
                zg_voice_destroy(v);
                (void)first;
                zg_graph* own = nullptr;
                if (zg_graph_compile("~(_2 + 0.5f*_1[_1]) |= _1 - _1[_3]", &own) != ZG_OK) bad[t]++;
                zg_graph_destroy(own);
                char msg[32];
                std::snprintf(msg, sizeof msg, "_1 |= (%d", t);                 // parse error, message names the thread's text
                zg_graph* none = nullptr;
                if (zg_graph_compile(msg, &none) == ZG_OK || !std::strstr(zg_last_error(), msg)) bad[t]++;
            }
        });
    for (auto& x : th) x.join();
    zg_graph_destroy(g);
    int total = 0;
    for (int b : bad) total += b;
    std::printf("8 threads x 200 rounds, %d failures\n", total);
    return total ? 1 : 0;
}
