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
