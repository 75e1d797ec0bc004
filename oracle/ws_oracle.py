"""CPU oracle of the MVS workspace helpers (TEST INFRASTRUCTURE: imported by tests/ only).

Independent pure-Python / numpy restatement of
  mvs::Model::ComputeDepthRanges / ComputeSharedPoints / ComputeTriangulationAngles / GetMaxOverlappingImages
      (src/colmap/mvs/model.cc:120-283), Percentile (src/colmap/math/math.h:205-224),
      CalculateTriangulationAngle (src/colmap/geometry/triangulation.cc:217-250)
  the map / consistency-graph file layouts (src/colmap/mvs/mat.cc:41-66, consistency_graph.cc:66-139)
written with dictionaries like the reference (not with the dense tables of the product).  Pinned to the reference's
own known answers (mvs/model_test.cc:84-196, mvs/consistency_graph_test.cc:40-138) in tests/test_workspace_cpu.py."""
import math

import numpy as np


def percentile(values, p):
    v = sorted(float(np.float32(x)) for x in values)
    idx = p / 100.0 * (len(v) - 1)
    lo, hi = int(math.floor(idx)), int(math.ceil(idx))
    if lo == hi:
        return v[hi]
    return (hi - idx) * v[lo] + (idx - lo) * v[hi]


def depth_ranges(images, points):
    """images: list of (R 3x3, T 3) float32; points: list of (xyz, track)."""
    depths = [[] for _ in images]
    for xyz, track in points:
        X = np.asarray(xyz, np.float32)
        for i in track:
            R, T = images[i]
            d = np.float32(np.float32(np.float32(R[2, 0] * X[0]) + np.float32(R[2, 1] * X[1])) + np.float32(R[2, 2] * X[2])) + np.float32(T[2])
            if d > 0:
                depths[i].append(np.float32(d))
    out = []
    for d in depths:
        if not d:
            out.append((-1.0, -1.0))
            continue
        d = sorted(d)
        lo = int(np.float32(len(d)) * np.float32(0.01))
        hi = int(np.float32(len(d)) * np.float32(0.99))
        out.append((float(np.float32(d[lo]) * np.float32(0.75)), float(np.float32(d[hi]) * np.float32(1.25))))
    return out


def shared_points(num_images, points):
    out = [dict() for _ in range(num_images)]
    for _, track in points:
        for i in range(len(track)):
            for j in range(i):
                a, b = track[i], track[j]
                if a != b:
                    out[a][b] = out[a].get(b, 0) + 1
                    out[b][a] = out[b].get(a, 0) + 1
    return out


def triangulation_angle(c1, c2, X):
    a, b = np.asarray(X, np.float64) - c1, np.asarray(X, np.float64) - c2
    n1, n2 = float(a @ a), float(b @ b)
    if n1 == 0.0 or n2 == 0.0:
        ang = 0.0
    else:
        ang = math.acos(min(1.0, max(-1.0, float(a @ b) / math.sqrt(n1 * n2))))
    return min(ang, math.pi - ang)


def triangulation_angles(images, points, p):
    centers = []
    for R, T in images:
        R = np.asarray(R, np.float32); T = np.asarray(T, np.float32)
        c = -(R.T @ T).astype(np.float32)       # fp32 like ComputeProjectionCenter
        centers.append(c.astype(np.float64))
    allang = [dict() for _ in images]
    for xyz, track in points:
        X = np.asarray(xyz, np.float32).astype(np.float64)
        for i in range(len(track)):
            for j in range(i):
                a, b = track[i], track[j]
                if a != b:
                    ang = np.float32(triangulation_angle(centers[a], centers[b], X))
                    allang[a].setdefault(b, []).append(ang)
                    allang[b].setdefault(a, []).append(ang)
    return [{k: float(np.float32(percentile(v, p))) for k, v in d.items()} for d in allang]


def max_overlapping_images(images, points, num, min_angle_deg):
    shared = shared_points(len(images), points)
    tri = triangulation_angles(images, points, 75.0)
    min_rad = float(np.float32(math.radians(min_angle_deg)))
    out = []
    for i in range(len(images)):
        cand = [(j, c) for j, c in sorted(shared[i].items()) if np.float32(tri[i][j]) >= np.float32(min_rad)]
        cand.sort(key=lambda t: -t[1])          # stable: ties keep ascending image index
        out.append([j for j, _ in cand[:num]])
    return out


def mat_bytes(arr):
    a = np.ascontiguousarray(arr, "<f4")
    if a.ndim == 2:
        a = a[None]
    d, h, w = a.shape
    return f"{w}&{h}&{d}&".encode() + a.tobytes()


def graph_bytes(width, height, data):
    return f"{width}&{height}&1&".encode() + np.ascontiguousarray(data, "<i4").tobytes()


# ---------------------------------------------------------------------------------------------------------------------
# stereo fusion (src/colmap/mvs/fusion.cc:131-545), single-threaded schedule, pure Python with numpy float32 scalars
# ---------------------------------------------------------------------------------------------------------------------
F = np.float32


def _median(values):
    return percentile(values, 50.0)


def _dot3(a0, a1, a2, b0, b1, b2):
    return F(F(F(a0 * b0) + F(a1 * b1)) + F(a2 * b2))


def fuse(opt, images, overlap):
    """opt: dict(min_num_pixels, max_num_pixels, max_traversal_depth, max_reproj_error, max_depth_error, max_normal_error,
    bbox_min, bbox_max); images: list of dict(used, K, R, T (float32), image_width, image_height, depth (h,w), normal
    (3,h,w), rgb (H,W,3), mask or None); overlap: list of lists.  Returns (xyz, normal, rgb, visibility)."""
    n = len(images)
    max_sq = F(opt["max_reproj_error"] * opt["max_reproj_error"])
    min_cos = F(math.cos(opt["max_normal_error"] * 0.017453292519943295))
    max_derr = F(opt["max_depth_error"])
    V, masks, used, fused = [None] * n, [None] * n, [False] * n, [False] * n
    for i, im in enumerate(images):
        if not im["used"]:
            continue
        used[i] = True
        h, w = im["depth"].shape
        sx, sy = F(F(w) / F(im["image_width"])), F(F(h) / F(im["image_height"]))
        K = np.array(im["K"], np.float32).reshape(3, 3).copy()
        K[0, 0] = F(K[0, 0] * sx); K[0, 2] = F(K[0, 2] * sx); K[1, 1] = F(K[1, 1] * sy); K[1, 2] = F(K[1, 2] * sy)
        R = np.asarray(im["R"], np.float32).reshape(3, 3); T = np.asarray(im["T"], np.float32).reshape(3)
        M = np.concatenate([R, T[:, None]], 1)
        P = np.empty((3, 4), np.float32)
        for r in range(3):
            for c in range(4):
                P[r, c] = _dot3(K[r, 0], K[r, 1], K[r, 2], M[0, c], M[1, c], M[2, c])
        Kd = K.astype(np.float64)
        det = Kd[0, 0] * (Kd[1, 1] * Kd[2, 2] - Kd[1, 2] * Kd[2, 1]) - Kd[0, 1] * (Kd[1, 0] * Kd[2, 2] - Kd[1, 2] * Kd[2, 0]) \
            + Kd[0, 2] * (Kd[1, 0] * Kd[2, 1] - Kd[1, 1] * Kd[2, 0])
        k = Kd.reshape(9)
        iK = np.array([(k[4] * k[8] - k[5] * k[7]) / det, (k[2] * k[7] - k[1] * k[8]) / det, (k[1] * k[5] - k[2] * k[4]) / det,
                       (k[5] * k[6] - k[3] * k[8]) / det, (k[0] * k[8] - k[2] * k[6]) / det, (k[2] * k[3] - k[0] * k[5]) / det,
                       (k[3] * k[7] - k[4] * k[6]) / det, (k[1] * k[6] - k[0] * k[7]) / det, (k[0] * k[4] - k[1] * k[3]) / det])
        Rd, Td = R.astype(np.float64), T.astype(np.float64)
        invP = np.empty((3, 4), np.float32)
        for r in range(3):
            for c in range(3):
                invP[r, c] = F((Rd[0, r] * iK[c] + Rd[1, r] * iK[3 + c]) + Rd[2, r] * iK[6 + c])
            invP[r, 3] = F(-((Rd[0, r] * Td[0] + Rd[1, r] * Td[1]) + Rd[2, r] * Td[2]))
        V[i] = dict(P=P, invP=invP, invR=R.T.copy(), sx=sx, sy=sy, w=w, h=h)
        masks[i] = np.zeros((h, w), bool) if im.get("mask") is None else (np.asarray(im["mask"]) != 0).copy()
    out_xyz, out_n, out_c, out_v = [], [], [], []

    def run(image0, row0, col0):
        queue = [(image0, row0, col0, 0)]
        ref_p, ref_n = np.zeros(4, np.float32), np.zeros(3, np.float32)
        px, py, pz, nx, ny, nz, cr, cg, cb, vis = [], [], [], [], [], [], [], [], [], set()
        while queue:
            i, row, col, td = queue.pop()
            im, v = images[i], V[i]
            if masks[i][row, col]:
                continue
            depth = F(im["depth"][row, col])
            if depth <= 0:
                continue
            P = v["P"]
            if td > 0:
                proj = [F(F(F(F(P[r, 0] * ref_p[0]) + F(P[r, 1] * ref_p[1])) + F(P[r, 2] * ref_p[2])) + F(P[r, 3] * ref_p[3])) for r in range(3)]
                if F(abs(F(F(proj[2] - depth) / depth))) > max_derr:
                    continue
                cd, rd = F(F(proj[0] / proj[2]) - F(col)), F(F(proj[1] / proj[2]) - F(row))
                if F(F(cd * cd) + F(rd * rd)) > max_sq:
                    continue
            nm = im["normal"]
            n0, n1, n2 = F(nm[0, row, col]), F(nm[1, row, col]), F(nm[2, row, col])
            iR = v["invR"]
            normal = [_dot3(iR[r, 0], iR[r, 1], iR[r, 2], n0, n1, n2) for r in range(3)]
            if td > 0 and _dot3(ref_n[0], ref_n[1], ref_n[2], normal[0], normal[1], normal[2]) < min_cos:
                continue
            hv = [F(F(col) * depth), F(F(row) * depth), depth, F(1)]
            iP = v["invP"]
            xyz = [F(F(F(F(iP[r, 0] * hv[0]) + F(iP[r, 1] * hv[1])) + F(iP[r, 2] * hv[2])) + F(iP[r, 3] * hv[3])) for r in range(3)]
            xx, yy = int(_round_half_away(float(F(F(col) / v["sx"])))), int(_round_half_away(float(F(F(row) / v["sy"]))))
            H, W = im["rgb"].shape[:2]
            color = tuple(int(c) for c in im["rgb"][yy, xx]) if 0 <= xx < W and 0 <= yy < H else (0, 0, 0)
            masks[i][row, col] = True
            if any(xyz[k] < F(opt["bbox_min"][k]) or xyz[k] > F(opt["bbox_max"][k]) for k in range(3)):
                continue
            px.append(xyz[0]); py.append(xyz[1]); pz.append(xyz[2]); nx.append(normal[0]); ny.append(normal[1]); nz.append(normal[2])
            cr.append(color[0]); cg.append(color[1]); cb.append(color[2]); vis.add(i)
            if td == 0:
                ref_p = np.array([xyz[0], xyz[1], xyz[2], 1], np.float32); ref_n = np.array(normal, np.float32)
            if len(px) >= opt["max_num_pixels"]:
                break
            if td >= opt["max_traversal_depth"] - 1:
                continue
            for nxt in overlap[i]:
                if not used[nxt] or fused[nxt]:
                    continue
                Pn = V[nxt]["P"]
                q = [F(F(F(F(Pn[r, 0] * xyz[0]) + F(Pn[r, 1] * xyz[1])) + F(Pn[r, 2] * xyz[2])) + Pn[r, 3]) for r in range(3)]
                nc, nr = int(_round_half_away(float(F(q[0] / q[2])))), int(_round_half_away(float(F(q[1] / q[2]))))
                if nc < 0 or nr < 0 or nc >= V[nxt]["w"] or nr >= V[nxt]["h"]:
                    continue
                queue.append((nxt, nr, nc, td + 1))
        if len(px) < opt["min_num_pixels"] or not px:
            return
        fn = [F(_median(nx)), F(_median(ny)), F(_median(nz))]
        norm = F(np.sqrt(F(F(F(fn[0] * fn[0]) + F(fn[1] * fn[1])) + F(fn[2] * fn[2]))))
        if norm < np.finfo(np.float32).eps:
            return
        out_xyz.append([F(_median(px)), F(_median(py)), F(_median(pz))])
        out_n.append([F(fn[0] / norm), F(fn[1] / norm), F(fn[2] / norm)])
        out_c.append([min(255, max(0, int(_round_half_away(float(F(_median(c))))))) for c in (cr, cg, cb)])
        out_v.append(sorted(vis))

    image = 0 if n else -1
    while image >= 0:
        if used[image]:
            for row in range(V[image]["h"]):
                for col in range(V[image]["w"]):
                    if not masks[image][row, col]:
                        run(image, row, col)
        fused[image] = True
        nxt = next((c for c in overlap[image] if used[c] and not fused[c]), -1)
        if nxt < 0:
            nxt = next((i for i in range(n) if used[i] and not fused[i]), -1)
        image = nxt
    return (np.array(out_xyz, np.float32).reshape(-1, 3), np.array(out_n, np.float32).reshape(-1, 3),
            np.array(out_c, np.uint8).reshape(-1, 3), out_v)


def _round_half_away(x):
    return math.floor(x + 0.5) if x >= 0 else -math.floor(-x + 0.5)
