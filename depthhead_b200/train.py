"""Hough-forest TRAINING on the GPU scorer.

Mirror of `HoughLearning` (src/hough/prediction.rs:38-234): the same constructor arguments, the same
sample extraction (`learn`, :145-226: sliding window, background test, truth from the head mask,
20 negatives + 20 positives per image), the same per-node callbacks of `HoughTreeFunctions`
(src/hough/houghforest.rs:196-311: `param_set`, `impurity`, `early_stop`, `comp_leaf_data`).

What is NOT the reference's: the tree-growing driver.  The reference hands the callbacks to
stamm 0.2.0's `train_forest_parallel`, whose source is not vendored, and draws its random numbers
from `rand::thread_rng()`; both are therefore builder-defined here and stated once:
  * every tree gets `size_of_subset_per_training` samples drawn without replacement (all of them
    if the set is smaller), kept in drawn order;
  * trees grow breadth-first, root at depth 0; candidates on which the reference's `impurity`
    itself aborts (an empty side: assert!(res.is_finite()); a determinant sum below -0.001:
    unreachable!()) are skipped; a node is a leaf iff `early_stop` says so or no candidate is left;
    otherwise the candidate with the SMALLEST impurity wins, the first of equals;
  * samples keep their order through a split; `children[bit]` receives the side `binarize` == bit;
  * a counter-based generator (splitmix64 of seed + k * golden ratio, `CounterRng`) stands in for
    thread_rng; the C++ trainer behind `dh_train_forest` consumes the same stream in the same
    order, so both grow identical forests from the same samples and seed.
The expensive part — `impurity` of every candidate of every node of a level — runs in
`train_score_kernel` through the C ABI (`dh_train_score_level`), bit-identical to the CPU
restatement of the callbacks (tests/test_gpu_train.py).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .api import Context, HoughPrediction, IntrinsicMatrix, default_context


class CounterRng:
    """draw k = splitmix64(seed + k * 0x9E3779B97F4A7C15), k = 1, 2, ..; uniform [0,1) = top 53 bits."""
    G = np.uint64(0x9E3779B97F4A7C15)

    def __init__(self, seed: int):
        self.seed = np.uint64(seed & 0xFFFFFFFFFFFFFFFF)
        self.k = 0

    def random(self, n: int) -> np.ndarray:
        with np.errstate(over="ignore"):
            ks = np.arange(self.k + 1, self.k + 1 + n, dtype=np.uint64)
            z = self.seed + ks * self.G
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
        self.k += n
        return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)

    def permutation(self, n: int) -> np.ndarray:
        """Fisher-Yates from the back: for i = n-1 .. 1 swap a[i], a[floor(u * (i + 1))]."""
        a = np.arange(n, dtype=np.int64)
        if n < 2:
            return a
        u = self.random(n - 1)
        js = (u * np.arange(n, 1, -1, dtype=np.float64)).astype(np.int64)
        for t, i in enumerate(range(n - 1, 0, -1)):
            j = js[t]
            a[i], a[j] = a[j], a[i]
        return a


class TrainSet:
    """Training samples resident on the GPU (dh_trainset)."""

    def __init__(self, patches, is_object, offsets, rotations, rw: int, rh: int, ctx: Context | None = None):
        self.ctx = ctx or default_context()
        p = np.ascontiguousarray(patches, np.uint16)
        if p.ndim != 3:
            raise ValueError("patches must be [n, sub_h, sub_w] uint16")
        self.n, self.sh, self.sw = p.shape
        self.rw, self.rh = int(rw), int(rh)
        self.is_object = np.ascontiguousarray(is_object, np.uint8).reshape(self.n)
        self.offsets = np.ascontiguousarray(offsets, np.float32).reshape(self.n, 3)
        self.rotations = np.ascontiguousarray(rotations, np.float64).reshape(self.n, 3)
        h = C.c_void_p()
        capi.check(capi.load().dh_trainset_create(self.ctx._h, capi.ptr(p), self.n, self.sw, self.sh, self.rw, self.rh,
                                                  capi.ptr(self.is_object), capi.ptr(self.offsets), capi.ptr(self.rotations),
                                                  C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            capi.load().dh_trainset_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def score_level(self, sample_idx, node_off, cand_rects, cand_thr, depth: int, steepness: float) -> np.ndarray:
        """impurity (houghforest.rs:250-295) of cand[k][j] on node k's samples -> SPLIT_DTYPE[n_nodes][m]."""
        idx = np.ascontiguousarray(sample_idx, np.uint32)
        off = np.ascontiguousarray(node_off, np.uint64)
        rects = np.ascontiguousarray(cand_rects, np.int32)
        thr = np.ascontiguousarray(cand_thr, np.float64)
        k = len(off) - 1
        m = thr.size // max(k, 1) if k else 0
        out = np.zeros((k, m), capi.SPLIT_DTYPE)
        capi.check(capi.load().dh_train_score_level(self.ctx._h, self._h, capi.ptr(idx), capi.ptr(off), k, capi.ptr(rects),
                                                    capi.ptr(thr), m, int(depth), float(steepness), capi.ptr(out)))
        return out

    def split_level(self, sample_idx, node_off, rects, thr) -> np.ndarray:
        """binarize (houghforest.rs:185-193) of ONE NodeParam per node over its samples -> bits u8."""
        idx = np.ascontiguousarray(sample_idx, np.uint32)
        off = np.ascontiguousarray(node_off, np.uint64)
        r = np.ascontiguousarray(rects, np.int32)
        t = np.ascontiguousarray(thr, np.float64)
        bits = np.zeros(int(off[-1]), np.uint8)
        capi.check(capi.load().dh_train_split_level(self.ctx._h, self._h, capi.ptr(idx), capi.ptr(off), len(off) - 1, capi.ptr(r),
                                                    capi.ptr(t), capi.ptr(bits)))
        return bits


def scale_and_replace(w: int, h: int, scale: float, rel_x, rel_y):
    """Rect::scale_and_replace (types.rs:82-91) of Rect(0, 0, w, h), vectorised: x0, y0, x1, y1."""
    if scale > 1.0:
        z = np.zeros_like(np.asarray(rel_x), np.int64)
        return z, z, z + w, z + h
    nw, nh = float(w) * scale, float(h) * scale
    nx = 0.0 + np.asarray(rel_x, np.float64) * (float(w) - nw)
    ny = 0.0 + np.asarray(rel_y, np.float64) * (float(h) - nh)
    x0, y0 = nx.astype(np.uint32).astype(np.int64), ny.astype(np.uint32).astype(np.int64)   # `as u32`: truncation
    return x0, y0, x0 + int(np.uint32(nw)), y0 + int(np.uint32(nh))


def _mat3_inv_f32(k: np.ndarray) -> np.ndarray:
    """Mat3::inv (meancov_estimation.rs:344-352) in f32: adjugate / det, element by element."""
    f = np.float32
    a, b, c, d, e, g, h_, i_, j = (f(k[0, 0]), f(k[0, 1]), f(k[0, 2]), f(k[1, 0]), f(k[1, 1]), f(k[1, 2]), f(k[2, 0]), f(k[2, 1]), f(k[2, 2]))
    det = f(f(a * f(f(e * j) - f(g * i_))) - f(d * f(f(b * j) - f(c * i_)))) + f(h_ * f(f(b * g) - f(c * e)))
    adj = np.array([[f(e * j) - f(g * i_), f(c * i_) - f(b * j), f(b * g) - f(c * e)],
                    [f(g * h_) - f(d * j), f(a * j) - f(c * h_), f(c * d) - f(a * g)],
                    [f(d * i_) - f(e * h_), f(b * h_) - f(a * i_), f(a * e) - f(b * d)]], np.float32)
    return (adj / det).astype(np.float32)


def extract_samples(depth: np.ndarray, mask: np.ndarray, intrinsic: IntrinsicMatrix, mid, rot, stepwidth: int, sub_w: int, sub_h: int):
    """The body of the per-image loop of HoughLearning::learn (prediction.rs:178-214): every
    non-background window with its truth.  Returns (origins [m,2] of the windows, is_object [m],
    offsets [m,3] f32, rotations [m,3] f64) in sliding-window order."""
    d = np.ascontiguousarray(depth, np.uint16)
    h, w = d.shape
    left_w, left_h = sub_w // 2, sub_h // 2
    right_w, right_h = sub_w - left_w, sub_h - left_h
    ys = np.arange(left_h, h - right_h, stepwidth)      # types.rs:371-383
    xs = np.arange(left_w, w - right_w, stepwidth)
    if len(ys) == 0 or len(xs) == 0:
        z = np.zeros((0, 3))
        return np.zeros((0, 2), np.int64), np.zeros(0, np.uint8), z.astype(np.float32), z
    # average_value_in_rect(whole window) > 0.0  <=>  any non-zero pixel (prediction.rs:189-190)
    nz = np.zeros((h + 1, w + 1), np.int64)
    nz[1:, 1:] = np.cumsum(np.cumsum(d != 0, 0), 1)
    X, Y = np.meshgrid(xs, ys)
    x0, y0 = X - left_w, Y - left_h
    cnt = nz[y0 + sub_h, x0 + sub_w] - nz[y0, x0 + sub_w] - nz[y0 + sub_h, x0] + nz[y0, x0]
    keep = cnt > 0
    X, Y, x0, y0 = X[keep], Y[keep], x0[keep], y0[keep]
    flag = np.asarray(mask)[Y, X] != 0                    # mask[(x, y)], prediction.rs:191
    z = d[Y, X].astype(np.float32)                         # depth[(x, y)] as f32
    # img_to_space_coord (types.rs:432-445), f32, products and sums unfused
    inv = _mat3_inv_f32(intrinsic.mat)
    xf, yf = X.astype(np.float32), Y.astype(np.float32)
    r = [(xf * inv[j, 0] + yf * inv[j, 1]) + np.float32(1.0) * inv[j, 2] for j in range(3)]
    c = z / r[2]
    p3 = np.stack([r[0] * c, r[1] * c, r[2] * c], 1).astype(np.float32)
    offs = (p3 - np.asarray(mid, np.float32)[None, :]).astype(np.float32)      # vec3 - mid, prediction.rs:198
    rots = np.repeat(np.asarray(rot, np.float32).astype(np.float64)[None, :], len(X), 0)  # rot[k] as f64, :176
    return np.stack([x0, y0], 1), flag.astype(np.uint8), offs, rots


class HoughLearning:
    """prediction.rs:38-234.  `learn` returns a HoughPrediction (and keeps the forest arrays in
    `self.last_forest`, the layout of synth.make_forest / dh_forest_from_arrays)."""

    def __init__(self, stepwidth: int, subimg_width: int, subimg_height: int, max_depth: int, num_of_trees: int,
                 subset_size_per_tree: int, subrect_feature_scale: float, feature_number_per_node: int,
                 min_subset_size_to_stop: int, steepness_weighting: float):
        s = float(subrect_feature_scale)
        # HoughTreeFunctions::new returns None for these (houghforest.rs:143-147)
        if s > 1.0 or s < 0.0 or feature_number_per_node == 0 or steepness_weighting <= 0.0:
            raise ValueError("Bad parameter for learning Houghforest")
        # random_subrect_iterator(...).unwrap() panics for a factor of 0 (types.rs:106-110)
        if s <= 0.0:
            raise ValueError("subrect_feature_scale must be in (0, 1]")
        self.stepwidth, self.sub_w, self.sub_h = int(stepwidth), int(subimg_width), int(subimg_height)
        self.max_depth, self.n_trees, self.subset = int(max_depth), int(num_of_trees), int(subset_size_per_tree)
        self.scale, self.n_features = s, int(feature_number_per_node)
        self.min_subset, self.steepness = int(min_subset_size_to_stop), float(steepness_weighting)
        self.last_forest = None

    # -- HoughTreeFunctions callbacks (host side)
    def param_set(self, rng, n: int):
        """param_set (houghforest.rs:227-246; types.rs:140-160): n random NodeParams -> rects [n,8],
        thresholds [n].  Five draws per candidate, in the reference's order: four offsets, threshold."""
        u = rng.random(5 * n).reshape(n, 5)
        a = scale_and_replace(self.sub_w, self.sub_h, self.scale, u[:, 0], u[:, 1])
        b = scale_and_replace(self.sub_w, self.sub_h, self.scale, u[:, 2], u[:, 3])
        rects = np.stack(list(a) + list(b), 1).astype(np.int32)
        thr = -256.0 + u[:, 4] * (256.0 - -256.0)
        return rects, thr

    def early_stop(self, depth: int, is_object: np.ndarray) -> bool:
        """early_stop (houghforest.rs:302-311)."""
        if not is_object.any():
            return True
        return depth >= self.max_depth or len(is_object) < self.min_subset

    @staticmethod
    def comp_leaf_data(ts: TrainSet, idx: np.ndarray):
        """comp_leaf_data (houghforest.rs:204-225): prob, offsets, rotations of the Object samples in set order."""
        obj = ts.is_object[idx] != 0
        pos = idx[obj]
        return float(len(pos)) / float(len(idx)), ts.offsets[pos], ts.rotations[pos]

    # -- builder-defined stand-in for stamm's trainer (see the module docstring)
    def train_tree(self, ts: TrainSet, subset: np.ndarray, rng) -> dict:
        rects, thr, child, leaves = [], [], [], []
        frontier = [(0, np.asarray(subset, np.uint32))]   # (node slot, samples); slot -1 = the tree is one leaf
        rects.append(None); thr.append(None); child.append([0, 0])
        root_is_leaf = False
        depth = 0
        while frontier:
            todo = []
            for slot, idx in frontier:
                if self.early_stop(depth, ts.is_object[idx]):
                    todo.append((slot, idx, None))
                else:
                    todo.append((slot, idx, self.param_set(rng, self.n_features)))
            split = [t for t in todo if t[2] is not None]
            decided = {}
            if split:
                node_off = np.concatenate([[0], np.cumsum([len(t[1]) for t in split])]).astype(np.uint64)
                all_idx = np.concatenate([t[1] for t in split])
                st = ts.score_level(all_idx, node_off, np.stack([t[2][0] for t in split]), np.stack([t[2][1] for t in split]),
                                    depth, self.steepness)
                chosen_r, chosen_t, chosen_k = [], [], []
                for k, t in enumerate(split):
                    imp = st[k]["impurity"]
                    ok = ~np.isnan(imp)
                    if ok.any():
                        j = int(np.flatnonzero(ok)[np.argmin(imp[ok])])      # smallest impurity, first of equals
                        chosen_r.append(t[2][0][j]); chosen_t.append(t[2][1][j]); chosen_k.append(k)
                if chosen_k:
                    sub_off = np.concatenate([[0], np.cumsum([len(split[k][1]) for k in chosen_k])]).astype(np.uint64)
                    sub_idx = np.concatenate([split[k][1] for k in chosen_k])
                    bits = ts.split_level(sub_idx, sub_off, np.stack(chosen_r), np.asarray(chosen_t))
                    for n_, k in enumerate(chosen_k):
                        b = bits[int(sub_off[n_]):int(sub_off[n_ + 1])]
                        decided[split[k][0]] = (chosen_r[n_], chosen_t[n_], split[k][1][b == 0], split[k][1][b != 0])
            nxt = []
            for slot, idx, _ in todo:
                if slot in decided:
                    r, t, left, right = decided[slot]
                    rects[slot], thr[slot] = r, t
                    for bit, part in ((0, left), (1, right)):
                        child[slot][bit] = len(rects)
                        rects.append(None); thr.append(None); child.append([0, 0])
                        nxt.append((child[slot][bit], part))
                else:
                    leaves.append((slot, self.comp_leaf_data(ts, idx)))
            frontier = nxt
            depth += 1
        # slots that became leaves are dropped from the node table; children are re-indexed
        leaf_of_slot = {slot: i for i, (slot, _) in enumerate(leaves)}
        node_slots = [s for s in range(len(rects)) if rects[s] is not None]
        node_of_slot = {s: i for i, s in enumerate(node_slots)}
        if 0 in leaf_of_slot:
            root_is_leaf = True
        out_child = []
        for s in node_slots:
            out_child.append([node_of_slot[c] if c in node_of_slot else ~leaf_of_slot[c] for c in child[s]])
        assert root_is_leaf or node_slots[0] == 0
        return dict(rects=np.asarray([rects[s] for s in node_slots], np.int32).reshape(-1, 8),
                    threshold=np.asarray([thr[s] for s in node_slots], np.float64),
                    child=np.asarray(out_child, np.int32).reshape(-1, 2), leaves=[l for _, l in leaves])

    def train_forest(self, ts: TrainSet, rng) -> dict:
        trees = []
        for _ in range(self.n_trees):
            subset = rng.permutation(ts.n)[:min(self.subset, ts.n)]   # drawn without replacement, in drawn order
            trees.append(self.train_tree(ts, subset, rng))
        node_off = np.concatenate([[0], np.cumsum([len(t["threshold"]) for t in trees])]).astype(np.int64)
        leaf_off = np.concatenate([[0], np.cumsum([len(t["leaves"]) for t in trees])]).astype(np.int64)
        leaves = [l for t in trees for l in t["leaves"]]
        vote_off = np.concatenate([[0], np.cumsum([len(l[1]) for l in leaves])]).astype(np.int64)
        return dict(n_trees=self.n_trees, tree_node_off=node_off, tree_leaf_off=leaf_off,
                    rects=np.concatenate([t["rects"] for t in trees]).astype(np.int64),
                    threshold=np.concatenate([t["threshold"] for t in trees]),
                    child=np.concatenate([t["child"] for t in trees]).astype(np.int32),
                    prob=np.asarray([l[0] for l in leaves], np.float64), vote_off=vote_off,
                    offsets=np.concatenate([l[1] for l in leaves]).astype(np.float32).reshape(-1, 3),
                    rotations=np.concatenate([l[2] for l in leaves]).astype(np.float64).reshape(-1, 3),
                    sub_w=self.sub_w, sub_h=self.sub_h, max_depth=self.max_depth)

    def train_native(self, gaussian_sigma: float, patches, is_object, offsets, rotations, seed: int,
                     ctx: Context | None = None) -> HoughPrediction:
        """The same tree growing in C++ behind the C ABI (dh_train_forest): identical forest for
        identical samples and seed."""
        ctx = ctx or default_context()
        p = capi.dh_train_params()
        p.stepwidth, p.subimage_width, p.subimage_height = self.stepwidth, self.sub_w, self.sub_h
        p.max_depth, p.n_trees, p.subset_per_tree = self.max_depth, self.n_trees, self.subset
        p.subrect_feature_scale, p.features_per_node, p.min_subset_size = self.scale, self.n_features, self.min_subset
        p.steepness, p.gaussian_sigma, p.seed = self.steepness, float(gaussian_sigma), int(seed) & 0xFFFFFFFFFFFFFFFF
        px = np.ascontiguousarray(patches, np.uint16)
        io = np.ascontiguousarray(is_object, np.uint8)
        of = np.ascontiguousarray(offsets, np.float32)
        ro = np.ascontiguousarray(rotations, np.float64)
        h = C.c_void_p()
        capi.check(capi.load().dh_train_forest(ctx._h, C.byref(p), capi.ptr(px), len(px), capi.ptr(io), capi.ptr(of), capi.ptr(ro),
                                               C.byref(h)))
        return HoughPrediction(h)

    def _params(self, gaussian_sigma: float, seed: int):
        p = capi.dh_train_params()
        p.stepwidth, p.subimage_width, p.subimage_height = self.stepwidth, self.sub_w, self.sub_h
        p.max_depth, p.n_trees, p.subset_per_tree = self.max_depth, self.n_trees, self.subset
        p.subrect_feature_scale, p.features_per_node, p.min_subset_size = self.scale, self.n_features, self.min_subset
        p.steepness, p.gaussian_sigma, p.seed = self.steepness, float(gaussian_sigma), int(seed) & 0xFFFFFFFFFFFFFFFF
        return p

    def learn_native(self, gaussian_sigma: float, data, seed: int = 0, ctx: Context | None = None) -> HoughPrediction:
        """HoughLearning::learn entirely in C++ behind the C ABI (dh_train_learn): sample extraction and
        tree growing; the same forest as learn() for the same frames and seed."""
        ctx = ctx or default_context()
        data = list(data)
        depth = np.ascontiguousarray(np.stack([np.asarray(t["depth"], np.uint16) for t in data]))
        mask = np.ascontiguousarray(np.stack([np.asarray(t["mask"], np.uint8) for t in data]))
        K = np.ascontiguousarray(np.stack([t["intrinsic"].mat.reshape(9) for t in data]), np.float32)
        pos = np.ascontiguousarray(np.stack([np.asarray(t["pos3d"], np.float32) for t in data]))
        rot = np.ascontiguousarray(np.stack([np.asarray(t["rot"], np.float32) for t in data]))
        n, h, w = depth.shape
        p = self._params(gaussian_sigma, seed)
        hnd = C.c_void_p()
        capi.check(capi.load().dh_train_learn(ctx._h, C.byref(p), n, w, h, capi.ptr(depth), capi.ptr(mask), capi.ptr(K), capi.ptr(pos),
                                              capi.ptr(rot), C.byref(hnd)))
        return HoughPrediction(hnd)

    def learn(self, gaussian_sigma: float, data, seed: int = 0, ctx: Context | None = None, native: bool = False) -> HoughPrediction:
        """HoughLearning::learn (prediction.rs:145-234).  `data`: iterable of dicts with `depth`
        [h,w] u16, `mask` [h,w] u8, `intrinsic` IntrinsicMatrix, `pos3d` [3], `rot` [3]
        (db_reader DepthTrue).  native=True grows the trees in C++ (dh_train_forest) instead of the
        Python loop; both give the same forest."""
        rng = CounterRng(seed)
        patches, is_obj, offs, rots = [], [], [], []
        for truth in data:
            org, flag, o, r = extract_samples(truth["depth"], truth["mask"], truth["intrinsic"], truth["pos3d"], truth["rot"],
                                              self.stepwidth, self.sub_w, self.sub_h)
            neg, pos = np.flatnonzero(flag == 0), np.flatnonzero(flag != 0)
            # rand_perm + take(20), negatives first (prediction.rs:216-226)
            for part in (neg[rng.permutation(len(neg))][:20], pos[rng.permutation(len(pos))][:20]):
                for i in part:
                    x0, y0 = org[i]
                    patches.append(np.asarray(truth["depth"])[y0:y0 + self.sub_h, x0:x0 + self.sub_w])   # to_cropped_subimage
                    is_obj.append(flag[i]); offs.append(o[i]); rots.append(r[i])
        if not patches:
            raise ValueError("no training samples")
        perm = rng.permutation(len(patches))            # rand_perm(train_ref), prediction.rs:229
        scale = self.scale
        rw, rh = int(np.uint32(float(self.sub_w) * scale)), int(np.uint32(float(self.sub_h) * scale))
        self.last_samples = (np.stack(patches)[perm], np.asarray(is_obj, np.uint8)[perm], np.asarray(offs, np.float32)[perm],
                             np.asarray(rots, np.float64)[perm])
        if native:
            self.last_forest = None
            return self.train_native(gaussian_sigma, *self.last_samples, seed=seed + 1, ctx=ctx)
        ts = TrainSet(*self.last_samples, rw, rh, ctx=ctx)
        try:
            self.last_forest = self.train_forest(ts, CounterRng(seed + 1))
        finally:
            ts.close()
        return HoughPrediction.from_arrays(self.last_forest, self.stepwidth, gaussian_sigma=gaussian_sigma, meanshift_iterations=20,
                                           subimage_width=self.sub_w, subimage_height=self.sub_h)
