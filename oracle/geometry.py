"""Oracle for the geometry half of the hot path.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Plain numpy / Python restatement of the reference's src/common/ code (wg-perception/tod 0.5.6); every function cites
the file:line it follows.  Neighbour lists are kept as sorted Python lists exactly like
tod::maximum_clique::AdjacencyMatrix (maximum_clique.h:52-148).  Validated in tests/test_oracle_geometry.py against
the reference's own sources compiled unmodified into oracle/_ref/libtod_ref.so (oracle/build_ref.py) and against the
two known-answer tests of test/test_maximum_clique.cpp.

Arithmetic notes: np.float32 scalars/arrays give IEEE single ops with one rounding each (no FMA), which is what the
reference's x86-64 build does; cv::norm(Vec3f) accumulates in double.
"""
import math

import numpy as np

F32 = np.float32


# ------------------------------------------------------------------------------------------------------------------
# sampler stream: restated copy of the product's tod_rng_* (tod_b200/csrc/common.cpp) — the harness makes the compiled
# reference's rand() (sac_model_registration_graph.h:111) return the same numbers.
# ------------------------------------------------------------------------------------------------------------------
_M64 = (1 << 64) - 1


def rng_seed(seed, object_index, rnd):
    z = (seed + 0x9E3779B97F4A7C15 * (object_index + 1) + 0xBF58476D1CE4E5B9 * (rnd + 1)) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


class Rng:
    def __init__(self, state):
        self.state = state & _M64

    def rand(self):
        self.state = (self.state * 6364136223846793005 + 1442695040888963407) & _M64
        return self.state >> 33


# ------------------------------------------------------------------------------------------------------------------
# FillAdjacency — adjacency_ransac.cpp:127-172
# ------------------------------------------------------------------------------------------------------------------
def fill_adjacency_dense(query, train, pixels, span, sensor_error):
    """Boolean n x n (physical, sample) matrices; entry (i, j) <=> j in neighbors(i)."""
    q = np.asarray(query, F32).reshape(-1, 3)
    t = np.asarray(train, F32).reshape(-1, 3)
    px = np.asarray(pixels, F32).reshape(-1, 2)
    n = q.shape[0]
    err = F32(sensor_error)
    e2 = F32(2) * err                      # `2 * sensor_error`           :144,163
    e4 = F32(4) * err                      # `4 * sensor_error`           :151
    sp = F32(span) + e2
    thr = sp * sp                          # (object_span + 2e)^2         :144
    with np.errstate(invalid="ignore", over="ignore"):
        d = q[:, None, :] - q[None, :, :]                                  # distSq, sac_model...h:52-58
        dq2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
        far = dq2 > thr                                                    # :144  continue
        dq = np.sqrt(dq2)                                                  # :146
        u = (t[:, None, :] - t[None, :, :]).astype(np.float64)             # Vec3f difference, then cv::norm in double
        s2 = (u[..., 0] * u[..., 0] + u[..., 1] * u[..., 1]) + u[..., 2] * u[..., 2]
        dt = np.sqrt(s2).astype(F32)                                       # :149  float dist_training = cv::norm(...)
        diff = np.abs(dt - dq)
        bad = diff > e4                                                    # :151  continue
        P = ~far & ~bad
        a = px[:, None, :] - px[None, :, :]
        pd = a[..., 0] * a[..., 0] + a[..., 1] * a[..., 1]                 # :158-160
        S = P & (pd > F32(400)) & (diff < e2)                              # :161-163
    np.fill_diagonal(P, False)
    np.fill_diagonal(S, False)
    assert (P == P.T).all() and (S == S.T).all()
    return P, S


def row_words(n):
    return ((n + 31) // 32 + 3) & ~3 if n > 0 else 0


def pack_bits(M):
    """Dense boolean n x n -> n x row_words(n) uint32 bit-matrix (bit j of row i)."""
    n = M.shape[0]
    W = row_words(n)
    out = np.zeros((n, W * 32), bool)
    out[:, :n] = M
    return np.packbits(out, axis=1, bitorder="little").view("<u4").reshape(n, W)


def unpack_bits(B, n):
    return np.unpackbits(np.ascontiguousarray(B).view(np.uint8).reshape(B.shape[0], -1), axis=1,
                         bitorder="little")[:, :n].astype(bool)


def dense_to_lists(M):
    return [list(np.nonzero(r)[0]) for r in M]


# ------------------------------------------------------------------------------------------------------------------
# Rigid fit — sac_model_registration_graph.h:304-347
# ------------------------------------------------------------------------------------------------------------------
def kabsch(query, train, indices):
    """estimateRigidTransformationSVD: returns (R 3x3 f32, T 3 f32) mapping query -> training frame."""
    q = np.asarray(query, F32).reshape(-1, 3)
    t = np.asarray(train, F32).reshape(-1, 3)
    ct = np.zeros(3, F32)
    cq = np.zeros(3, F32)
    for i in indices:                              # :312-315 float accumulation in index order
        ct = ct + t[i]
        cq = cq + q[i]
    inv = F32(1.0) / F32(len(indices))             # cv::Vec operator/= multiplies by 1.f/alpha
    ct = ct * inv
    cq = cq * inv
    st = (t[list(indices)] - ct).astype(F32)       # :320-326
    sq = (q[list(indices)] - cq).astype(F32)
    H = (st.astype(np.float64).T @ sq.astype(np.float64)).astype(F32)     # :330 gemm, f64 accumulate
    U, _, Vt = np.linalg.svd(H.astype(np.float64))                        # :333 cv::SVD (float Jacobi in OpenCV)
    if np.linalg.det(U) * np.linalg.det(Vt) < 0:                          # :337-341
        Vt = Vt.copy()
        Vt[2, :] *= -1
    R = (U @ Vt).astype(F32)                                              # :343
    T = (ct - (R @ cq).astype(F32)).astype(F32)                           # :344
    return R, T


def transform_dist_sq(R, T, q, t):
    """distSq(R * pt_src + T, pt_tgt) in float (sac_model...h:198)."""
    R = np.asarray(R, F32)
    p = np.empty(3, F32)
    for r in range(3):
        p[r] = F32(F32(F32(R[r, 0] * q[0]) + F32(R[r, 1] * q[1])) + F32(R[r, 2] * q[2])) + T[r]
    d = p - t
    return F32(F32(d[0] * d[0] + d[1] * d[1]) + d[2] * d[2])


# ------------------------------------------------------------------------------------------------------------------
# maximum_clique::Graph — maximum_clique.cpp:209-375
# ------------------------------------------------------------------------------------------------------------------
class Graph:
    def __init__(self, n):
        self.adj = [[] for _ in range(n)]  # sorted neighbour lists

    def add_edge(self, a, b):             # AdjacencyMatrix::set -> SetOneWay, maximum_clique.h:120-145
        for x, y in ((a, b), (b, a)):
            row = self.adj[x]
            import bisect
            k = bisect.bisect_left(row, y)
            if k == len(row) or row[k] != y:
                row.insert(k, y)

    def add_edge_sorted(self, a, b):      # set_sorted, maximum_clique.h:85-90
        self.adj[a].append(b)
        self.adj[b].append(a)

    def delete_edge(self, a, b):          # invalidate(i1, i2) -> InvalidateOneWay, maximum_clique.h:111-118
        self.adj[a].remove(b)
        self.adj[b].remove(a)

    def test(self, i, j):                 # binary_search, maximum_clique.cpp:122-126
        import bisect
        row = self.adj[i]
        k = bisect.bisect_left(row, j)
        return k < len(row) and row[k] == j

    # -- maximum_clique.cpp:263-284
    def degree_sort(self, R):
        n = len(R)
        deg = [[0, R[i]] for i in range(n)]
        for i in range(n):
            for j in range(i):
                if self.test(R[i], R[j]):
                    deg[i][0] += 1
                    deg[j][0] += 1
        deg.sort()
        for i in range(n):
            R[i] = deg[n - 1 - i][1]

    # -- maximum_clique.cpp:219-261.  C is the ONE colour vector shared by every recursion level, exactly as in the
    # reference (it is passed by reference all the way down).  It is modelled as a fixed-capacity array plus a size:
    # ColorSort writes C[0..|R|) without touching the size, MaxCliqueDyn reads C.back() and pops.  Once the size has
    # run to zero the reference reads/pops OUT OF BOUNDS (undefined behaviour).  What the compiled reference then
    # sees on glibc/x86-64 is the malloc chunk header in front of the array: data[-1] = high word of the chunk size
    # (0), data[-2] = low word of the chunk size | PREV_INUSE; below that, unknowable memory (taken as 0 here).
    # See DESIGN.md "clique quirks".  The gate (minimal_size 7) never depends on it in 300/300 random graphs.
    def color_sort(self, R, C, QMax, Q):
        min_k = max(1, len(QMax) - len(Q) + 1)
        Ck = [[], []]
        j = 0
        maxno = 2
        for p in list(R):
            k = 1
            while any(self.test(p, v) for v in Ck[k]):
                k += 1
                if k >= maxno:
                    maxno += 1
                    Ck.append([])
                    break
            if k < min_k:
                R[j] = p
                j += 1
            else:
                Ck[k].append(p)
        data = C[0]
        if j > 0:
            data[j - 1] = 0
        pos = j
        for k in range(min_k, maxno):
            for v in Ck[k]:
                R[pos] = v
                data[pos] = k
                pos += 1

    # -- maximum_clique.cpp:286-336
    def _max_clique_dyn(self, R, C, level, minimal_size, QMax, Q, S, SOld):
        if len(QMax) >= minimal_size:
            return
        if level >= len(S):
            S.append(0)
            SOld.append(0)
        S[level] = S[level] + S[level - 1] - SOld[level]
        SOld[level] = S[level - 1]
        while R:
            p = R[-1]
            if C[1] > 0:                                      # C.back()
                c = C[0][C[1] - 1]
            elif C[1] == -1:
                c = (max(32, (4 * len(C[0]) + 8 + 15) & ~15)) | 1
            else:
                c = 0
            if len(Q) + c > len(QMax):
                Q.append(p)
                Rp = [v for v in R if self.test(p, v)]        # Intersection, :209-217
                if Rp:
                    if S[level] / self.all_steps < self.t_limit:
                        self.degree_sort(Rp)
                    self.color_sort(Rp, C, QMax, Q)
                    S[level] += 1
                    self.all_steps += 1
                    if self.all_steps > 100000:
                        return
                    self._max_clique_dyn(Rp, C, level + 1, minimal_size, QMax, Q, S, SOld)
                elif len(Q) > len(QMax):
                    QMax[:] = Q
                    if len(QMax) >= minimal_size:
                        return
                Q.pop()
            else:
                return
            R.pop()
            C[1] -= 1                                         # C.pop_back() (may run negative, see above)

    # -- maximum_clique.cpp:343-369
    def find_clique(self, minimal_size):
        n = len(self.adj)
        QMax = []
        if n == 0:
            return QMax
        self.all_steps = 1
        self.t_limit = 0.025
        R = list(range(n))
        self.degree_sort(R)
        max_degree = len(self.adj[R[0]])
        data = [0] * n
        for i in range(min(max_degree, n)):
            data[i] = i + 1
        for i in range(max_degree, n):
            data[i] = max_degree + 1
        C = [data, n]                                         # [storage, size]
        S = [0] * (n + 1)
        SOld = [0] * (n + 1)
        self._max_clique_dyn(R, C, 1, minimal_size, QMax, [], S, SOld)
        return QMax

    def find_maximum_clique(self):            # :371-375
        return self.find_clique(0xFFFFFFFF)


# ------------------------------------------------------------------------------------------------------------------
# AdjacencyRansac — adjacency_ransac.{h,cpp}
# ------------------------------------------------------------------------------------------------------------------
def _intersect_sorted(a, b):
    sb = set(b)
    return [x for x in a if x in sb]


class AdjacencyRansac:
    def __init__(self):
        self.training_points = []
        self.query_points = []
        self.query_indices = []
        self.valid_indices = []
        self.physical = []
        self.sample = []
        self.min_sample_size = 3

    def add_points(self, training_point, query_point, query_index):      # adjacency_ransac.cpp:51-59
        self.valid_indices.append(len(self.query_indices))
        self.training_points.append(np.asarray(training_point, F32))
        self.query_points.append(np.asarray(query_point, F32))
        self.query_indices.append(int(query_index))

    def fill_adjacency(self, keypoints_xy, span, sensor_error):          # :127-172
        kp = np.asarray(keypoints_xy, F32).reshape(-1, 2)
        px = kp[self.query_indices] if self.query_indices else np.zeros((0, 2), F32)
        P, S = fill_adjacency_dense(np.array(self.query_points, F32).reshape(-1, 3),
                                    np.array(self.training_points, F32).reshape(-1, 3), px, span, sensor_error)
        self.physical = dense_to_lists(P)
        self.sample = dense_to_lists(S)
        self.invalidate_indices([])                                       # :170-171 (a no-op: the loop never runs)

    @staticmethod
    def _invalidate_cluster(adj, indices):                                # maximum_clique.cpp:59-86
        s = set(indices)
        done = set()
        for index in indices:
            adj[index] = [v for v in adj[index] if v not in s]
            for sub in adj[index]:
                if sub in done:
                    continue
                adj[sub] = [v for v in adj[sub] if v not in s]
                done.add(sub)
            adj[index] = []

    def invalidate_indices(self, indices):                                # adjacency_ransac.cpp:63-89
        todo = list(indices)
        while todo:
            todo = sorted(set(todo))
            s = set(todo)
            self.valid_indices = [v for v in self.valid_indices if v not in s]
            self._invalidate_cluster(self.physical, todo)
            self._invalidate_cluster(self.sample, todo)
            todo = [v for v in self.valid_indices if len(self.sample[v]) < self.min_sample_size]

    def invalidate_query_indices(self, query_indices):                    # :93-123 (guarded against the end deref, Q11)
        if not len(query_indices):
            return
        qs = sorted(set(int(x) for x in query_indices))
        end = len(qs)
        it = 0
        to_remove = []
        for index in self.valid_indices:
            query_index = self.query_indices[index]
            if it < end and query_index < qs[it]:
                continue
            while it != end and query_index > qs[it]:
                it += 1
            if it != end and query_index == qs[it]:
                to_remove.append(index)
                continue
            if it == end:
                break
        self.invalidate_indices(to_remove)

    # -- the RANSAC model: sac_model_registration_graph.h ------------------------------------------------------------
    def _draw_helper(self, valid_samples, n_samples, rng, out):           # :102-132
        if n_samples == 0:
            return True
        if not valid_samples:
            return False
        while True:
            sample = valid_samples[rng.rand() % len(valid_samples)]
            new_valid = _intersect_sorted(valid_samples, self.sample[sample])
            if self._draw_helper(new_valid, n_samples - 1, rng, out):
                out.append(sample)
                return True
            valid_samples.remove(sample)
            if not valid_samples:
                return False

    def get_samples(self, rng, max_sample_checks=1000):                   # :141-168
        if len(self.valid_indices) < 3:
            return []
        for _ in range(max_sample_checks):
            out = []
            if self._draw_helper(list(self.valid_indices), 3, rng, out):
                return out
        return []

    def candidates(self, samples):                                        # :178-186
        poss = list(self.physical[samples[0]])
        for s in samples[1:]:
            poss = _intersect_sorted(poss, self.physical[s])
        return poss + list(samples)

    def select_within_distance(self, samples, R, T, threshold, state):    # :171-269; state = {"best": 8}
        if not samples:
            return []
        poss = self.candidates(samples)
        thr2 = float(threshold) * float(threshold) if threshold < 1e150 else math.inf
        inliers = []
        for i in poss:                                                    # :192-200
            d2 = float(transform_dist_sq(R, T, self.query_points[i], self.training_points[i]))
            if d2 < thr2:
                inliers.append(i)
        minimal_size = min(state["best"], 7)                              # :203
        if len(inliers) <= minimal_size:
            return inliers
        filtered = [v for v in inliers if len(self.sample[v]) >= minimal_size]   # :209-213
        if len(filtered) <= minimal_size:
            return []
        filtered.sort()
        max_possible = 0
        for v in filtered:                                                # :222-233
            max_possible = len(_intersect_sorted(self.sample[v], filtered))
            if max_possible > minimal_size:
                break
        if max_possible <= minimal_size:
            return []
        index = {v: j for j, v in enumerate(filtered)}                    # :241-255
        g = Graph(len(filtered))
        for j in range(len(filtered) - 1):
            tail = set(filtered[j + 1:])
            for nb in self.sample[filtered[j]]:
                if nb in tail:
                    g.add_edge_sorted(j, index[nb])
        vertices = g.find_clique(minimal_size)                            # :258-265
        if len(vertices) <= minimal_size:
            return []
        inliers.sort()
        state["best"] = max(len(inliers), state["best"])                  # :267-268
        return inliers

    def compute_model(self, rng, max_iterations, threshold=math.inf, trace=None):   # ransac.h:80-143
        iterations = 0
        n_best = -(2 ** 31 - 1)
        k = 1.0
        best_inliers, best_R, best_T = [], None, None
        state = {"best": 8}
        eps = np.finfo(np.float64).eps
        while iterations < k:
            selection = self.get_samples(rng)
            if not selection:
                break
            R, T = kabsch(self.query_points, self.training_points, selection)
            inliers = self.select_within_distance(selection, R, T, threshold, state)
            if trace is not None:
                trace.append((tuple(selection), len(inliers)))
            if len(inliers) > n_best:
                n_best = len(inliers)
                best_inliers, best_R, best_T = list(inliers), R, T
                w = n_best / float(len(self.valid_indices))
                p_no = 1.0 - w ** 3
                p_no = max(eps, p_no)
                p_no = min(1.0 - eps, p_no)
                k = math.log(1.0 - 0.99) / math.log(p_no)
            iterations += 1
            if iterations > max_iterations:
                break
        return best_inliers, best_R, best_T

    def ransac(self, sensor_error, n_ransac_iterations, rng, threshold=math.inf):   # adjacency_ransac.cpp:234-309
        """Returns (inlier keypoint indices sorted unique, R 3x3 f32 object->camera, T 3 f32), or ([], None, None)."""
        if len(self.valid_indices) < 3:
            return [], None, None
        inliers, R, T = self.compute_model(rng, n_ransac_iterations, threshold)
        if not inliers:
            return [], None, None
        inliers = sorted(inliers)
        s = set(inliers)
        valid = [v for v in self.valid_indices if v not in s]
        do_final = False
        thresh = float(F32(sensor_error) * F32(sensor_error))             # double thresh = float * float  :267
        while True:
            R, T = kabsch(self.query_points, self.training_points, inliers)   # :272
            extra = []
            for i in valid:                                                # :276-283
                q, t = self.query_points[i], self.training_points[i]
                p = np.empty(3, F32)
                for r in range(3):
                    p[r] = F32(F32(F32(R[r, 0] * q[0]) + F32(R[r, 1] * q[1])) + F32(R[r, 2] * q[2])) + T[r]
                d = (p - t).astype(np.float64)
                nrm = math.sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2])   # cv::norm -> double
                if nrm * nrm < thresh:
                    extra.append(i)
            inliers = sorted(inliers + extra)                              # std::merge of two sorted lists
            es = set(extra)
            valid = [v for v in valid if v not in es]
            if do_final:
                break
            if not extra:
                do_final = True
                thresh *= 4
        Rt = np.ascontiguousarray(R.T)                                     # :304
        T2 = np.empty(3, F32)
        for r in range(3):                                                 # T = -R * T  (R already transposed) :305
            T2[r] = F32(F32(F32((-Rt[r, 0]) * T[0]) + F32((-Rt[r, 1]) * T[1])) + F32((-Rt[r, 2]) * T[2]))
        kp = sorted(set(self.query_indices[i] for i in inliers))           # :306-308
        return kp, Rt, T2


def cluster_per_object(keypoints_xy, cloud, matches, counts, points3d):     # adjacency_ransac.cpp:176-205
    """matches: structured [nq,k]; returns dict imgIdx -> AdjacencyRansac (insertion in ascending query order)."""
    objs = {}
    kp = np.asarray(keypoints_xy, F32).reshape(-1, 2)
    for qi in range(matches.shape[0]):
        x, y = int(kp[qi, 0]), int(kp[qi, 1])          # at<Vec3f>(pt.y, pt.x): float -> int truncation (Q9)
        qp = cloud[y, x]
        if np.isnan(qp[0]):                            # cvIsNaN(query_point[0]) :189
            continue
        for j in range(int(counts[qi])):
            o = int(matches["imgIdx"][qi, j])
            objs.setdefault(o, AdjacencyRansac()).add_points(points3d[qi, j], qp, qi)
    return objs


def guess_process(keypoints_xy, cloud, matches, counts, points3d, spans, min_inliers, n_ransac_iterations,
                  sensor_error, seed, threshold=math.inf):                 # GuessGenerator.cpp:127-250
    """Returns list of (object_index, R, T, inlier_keypoints)."""
    poses = []
    objs = cluster_per_object(keypoints_xy, cloud, matches, counts, points3d)
    for o in sorted(objs):                                                  # std::map order :170-176
        ar = objs[o]
        ar.fill_adjacency(keypoints_xy, spans[o], sensor_error)             # :189
        rnd = 0
        while True:                                                         # :192-231
            rng = Rng(rng_seed(seed, o, rnd))
            inl, R, T = ar.ransac(sensor_error, n_ransac_iterations, rng, threshold)
            rnd += 1
            if len(inl) < min_inliers:
                break
            ar.invalidate_query_indices(inl)
            poses.append((o, R, T, inl))
    return poses
