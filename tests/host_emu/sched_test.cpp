// tests/host_emu/sched_test.cpp -- drives mdz_b200/csrc/band_grants.h (the host-side band scheduler's
// policy) with simulated devices: every band handed out exactly once, and the time the slowest device
// finishes stays close to total work / total speed although the devices differ in speed and the bands in
// cost.  usage: sched_test TOTAL_BANDS LOW_WATER MIN_CHUNK speed... ; band cost: 1, or 12 for the bands
// between 30 % and 45 % of the image (a dense interior).  Prints "OK makespan ideal grants".
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../../mdz_b200/csrc/band_grants.h"

int main(int argc, char** argv)
{
    if (argc < 5) return 2;
    const int total = atoi(argv[1]), low = atoi(argv[2]), minc = atoi(argv[3]);
    std::vector<double> speed;
    for (int i = 4; i < argc; ++i) speed.push_back(atof(argv[i]));
    const int n = (int)speed.size();
    auto cost = [&](int b) { return (b >= total * 30 / 100 && b < total * 45 / 100) ? 12.0 : 1.0; };
    mdz::BandGrants g(total, n, minc);
    std::vector<std::vector<int> > queue(n);          // bands granted, not yet started
    std::vector<double> busy_until(n, 0.0), left(n, 0.0);
    std::vector<int> seen(total, 0);
    std::vector<double> started(n, 0.0);
    double now = 0.0, work = 0.0, sp = 0.0;
    for (int b = 0; b < total; ++b) work += cost(b);
    for (int i = 0; i < n; ++i) sp += speed[i];
    int grants = 0;
    const double dt = 0.05;                            // the host's polling period, in units of one cheap band at speed 1
    for (;;) {
        bool any = false;
        for (int i = 0; i < n; ++i) {
            while ((int)queue[i].size() < low && g.remaining() > 0) {
                int first, cnt = g.take(i, &first);
                for (int k = 0; k < cnt; ++k) { if (seen[first + k]++) { printf("FAIL band %d twice\n", first + k); return 1; } queue[i].push_back(first + k); }
                ++grants;
            }
            // consume for dt
            double budget = dt * speed[i];
            while (budget > 0) {
                if (left[i] <= 0) {
                    if (queue[i].empty()) break;
                    left[i] = cost(queue[i].front()); queue[i].erase(queue[i].begin());
                    started[i] += 1.0; g.progress(i, started[i]);
                }
                const double use = left[i] < budget ? left[i] : budget;
                left[i] -= use; budget -= use;
                busy_until[i] = now + dt;
            }
            if (left[i] > 0 || !queue[i].empty()) any = true;
        }
        now += dt;
        if (!any && g.remaining() == 0) break;
        if (now > 1e7) { puts("FAIL no end"); return 1; }
    }
    for (int b = 0; b < total; ++b) if (seen[b] != 1) { printf("FAIL band %d seen %d\n", b, seen[b]); return 1; }
    double makespan = 0;
    for (int i = 0; i < n; ++i) if (busy_until[i] > makespan) makespan = busy_until[i];
    printf("OK %.2f %.2f %d\n", makespan, work / sp, grants);
    return 0;
}
