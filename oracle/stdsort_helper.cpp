// TEST INFRASTRUCTURE ONLY (oracle/).  Never linked into or called from the product path.
//
// detectGridFAST keeps the first element of a libstdc++ std::sort() by descending response
// (/root/reference/src/feature_extractor.cpp:75-77,518-521).  std::sort is not stable, so with
// equal maxima the winner depends on libstdc++'s introsort.  This helper calls the real
// std::sort (the same library the reference is compiled against with gcc) so the Python oracle
// and the CUDA restatement of introsort can both be checked against the genuine article.
#include <algorithm>
#include <vector>

namespace {
struct Kp { float response; int idx; };
bool compare_response(Kp a, Kp b) { return a.response > b.response; }
}

extern "C" void ov2_oracle_sort_desc(const float* resp, int n, int* order_out) {
    std::vector<Kp> v(n);
    for (int i = 0; i < n; ++i) { v[i].response = resp[i]; v[i].idx = i; }
    std::sort(v.begin(), v.end(), compare_response);
    for (int i = 0; i < n; ++i) order_out[i] = v[i].idx;
}
