#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include "../include/cgs_b200.h"
int main(int argc, char **argv) {
    for (int i = 1; i < argc; ++i) {
        for (int T : {1, 3, 8}) {
            cgs_bed_table *t = nullptr;
            int rc = cgs_bed_read(&t, argv[i], 3, T);
            if (rc) { printf("%s T=%d error: %s\n", argv[i], T, cgs_last_error()); continue; }
            int64_t n, cc, nc; int32_t cols, nch, nn, st;
            cgs_bed_dims(t, &n, &cols, &nch, &cc, &nn, &nc, &st);
            int64_t removed = -1;
            if (cols == 4) cgs_bed_collapse_duplicates(t, T, &removed);
            cgs_bed_dims(t, &n, &cols, &nch, &cc, &nn, &nc, &st);
            std::vector<int32_t> a(n), b(n), c(n), d(n), e(n);
            std::vector<char> s1(cc + 1), s2(nc + 1); std::vector<int64_t> o1(nch + 1), o2(nn + 1);
            rc = cgs_bed_export(t, a.data(), b.data(), c.data(), cols >= 4 ? d.data() : nullptr, st ? e.data() : nullptr, s1.data(), o1.data(), s2.data(), o2.data());
            long long sum = 0; for (auto v : b) sum += v;
            printf("%s T=%d rows=%lld cols=%d chrom=%d names=%d score=%d removed=%lld sum=%lld rc=%d\n", argv[i], T, (long long)n, cols, nch, nn, st, (long long)removed, sum, rc);
            cgs_bed_destroy(t);
        }
    }
}
