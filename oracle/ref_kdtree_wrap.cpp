// oracle/ref_kdtree_wrap.cpp -- thin C wrapper around the REFERENCE's own vendored nanoflann
// (pgo/SC-A-LOAM/include/scancontext/{nanoflann.hpp,KDTreeVectorOfVectorsAdaptor.h}).
//
// TEST INFRASTRUCTURE ONLY.  Built by oracle/Makefile into oracle/_ref/ (git-ignored) from the
// headers where they lie under /root/reference; no reference source is copied into this repo.
// It reproduces the candidate step of SCManager::detectLoopClosureID
// (SC/Scancontext.cpp:352-356 tree build, :367-373 kNN(3)) with the real kd-tree so that the
// oracle's brute-force restatement of the metric (oracle/sc_ref.c sc_knn_bruteforce) and the
// product's ring-key search can be pinned against the reference's candidate set.
#include <cstddef>
#include <memory>
#include <vector>

#include "scancontext/nanoflann.hpp"
#include "scancontext/KDTreeVectorOfVectorsAdaptor.h"

using KeyMat = std::vector<std::vector<float>>;                    // Scancontext.h:41
using InvKeyTree = KDTreeVectorOfVectorsAdaptor<KeyMat, float>;    // Scancontext.h:42

extern "C" int ref_kdtree_knn(const float *keys, int n, int dim, const float *query, int k,
                              int *out_idx, float *out_d2)
{
    KeyMat mat(n, std::vector<float>(dim));
    for (int i = 0; i < n; ++i)
        for (int d = 0; d < dim; ++d) mat[i][d] = keys[(size_t)i * dim + d];
    // Scancontext.cpp:356: InvKeyTree(PC_NUM_RING /* dim */, keys_to_search, 10 /* max leaf */)
    std::unique_ptr<InvKeyTree> tree = std::make_unique<InvKeyTree>(dim, mat, 10);
    std::vector<size_t> idx(k);
    std::vector<float> d2(k);
    nanoflann::KNNResultSet<float> result(k);                        // Scancontext.cpp:371-373
    result.init(&idx[0], &d2[0]);
    tree->index->findNeighbors(result, query, nanoflann::SearchParams(10));
    for (int j = 0; j < k; ++j) { out_idx[j] = (int)idx[j]; out_d2[j] = d2[j]; }
    return (int)result.size();
}
