"""Scene ingest without third-party packages.

Mirror of the reference ``doper/utils/assets.py`` (``line_begin_end`` :10-25, ``sort_segments``
:28-82, ``get_svg_scene`` :85-131) that needs neither ``svgpathtools`` nor ``tripy`` (both absent
from this image, as are the reference's map assets: every ``assets/**/*.svg`` is a git-lfs stub).

  * ``read_svg_paths``  — the subset of SVG the reference accepts: ``<svg width height>`` and
    ``<path d=...>`` made of straight lines (M/m, L/l, H/h, V/v, Z/z), optional
    ``transform="rotate(angle, cx, cy)"``.  ``scripts/generate_maps.py`` (svgpathtools ``wsvg``)
    writes exactly ``d="M x,y L x,y ..."``.
  * ``sort_segments``   — restatement of the reference's ordering (decides vertex order).
  * ``earclip``         — restatement of ear clipping as the ``tripy`` package implements it
    (``tripy.earclip``; the reference leaves tripy UNPINNED, ``setup.py:18``, call site
    ``assets.py:116``).  It decides the triangle order and therefore the SEGMENT INDICES of a
    scene.  PARITY UNPINNED: tripy is not installed, the reference has no test or golden scene,
    so index-level agreement with a reference-loaded map cannot be checked here; the tests check
    the invariants the kernels need (CCW triangles, 3 segments each, polygon-major, area
    preserved).
  * ``save_scene`` / ``load_scene`` — on-disk cache of the ``[S,2,2]`` float32 segments.

Load-time host code (float64 geometry like the reference's numpy), not part of any measured path.
"""
from __future__ import annotations

import math
import re
import sys
from pathlib import Path
from typing import List, Sequence, Tuple

import numpy as onp

from ..sim.jax_geometry import JaxScene, Polygon, create_polygon, create_scene

__all__ = ["line_begin_end", "sort_segments", "earclip", "read_svg_paths", "get_svg_scene", "rotate_polygon",
           "save_scene", "load_scene", "write_svg"]

Point = Tuple[float, float]


# ------------------------------------------------------------------------------------------
# SVG subset reader / writer
# ------------------------------------------------------------------------------------------
_NUM = r"[-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?"


def _attr(tag: str, name: str):
    m = re.search(r'\b' + re.escape(name) + r'\s*=\s*"([^"]*)"', tag) or \
        re.search(r'\b' + re.escape(name) + r"\s*=\s*'([^']*)'", tag)
    return m.group(1) if m else None


def _path_lines(d: str) -> List[Tuple[complex, complex]]:
    """Straight-line path data -> list of (start, end) complex pairs; raises on curves."""
    tokens = re.findall(r"[A-Za-z]|" + _NUM, d)
    lines: List[Tuple[complex, complex]] = []
    cur = start = 0j
    cmd, i = None, 0

    def num():
        nonlocal i
        v = float(tokens[i])
        i += 1
        return v

    while i < len(tokens):
        if re.fullmatch(r"[A-Za-z]", tokens[i]):
            cmd = tokens[i]
            i += 1
            if cmd in "Zz":
                if cur != start:
                    lines.append((cur, start))
                cur = start
                continue
        if cmd is None:
            raise ValueError("path data must start with a command")
        if cmd in "Mm":
            p = complex(num(), num())
            cur = start = (cur + p) if cmd == "m" else p  # a leading relative moveto is absolute (cur = 0)
            cmd = "L" if cmd == "M" else "l"  # subsequent pairs are implicit line-tos
        elif cmd in "Ll":
            p = complex(num(), num())
            nxt = cur + p if cmd == "l" else p
            lines.append((cur, nxt))
            cur = nxt
        elif cmd in "Hh":
            x = num()
            nxt = complex(cur.real + x, cur.imag) if cmd == "h" else complex(x, cur.imag)
            lines.append((cur, nxt))
            cur = nxt
        elif cmd in "Vv":
            y = num()
            nxt = complex(cur.real, cur.imag + y) if cmd == "v" else complex(cur.real, y)
            lines.append((cur, nxt))
            cur = nxt
        else:
            raise ValueError(f"only straight-line path commands are supported, got {cmd!r}")
    return lines


def read_svg_paths(fname) -> Tuple[List[List[Tuple[complex, complex]]], List[dict], dict]:
    """-> (paths as lists of (start, end), per-path attributes, svg attributes) — the three values
    the reference takes from ``svg2paths(fname, return_svg_attributes=True)`` (assets.py:96)."""
    text = Path(fname).read_text()
    svg_tag = re.search(r"<svg\b[^>]*>", text, re.S)
    if not svg_tag:
        raise ValueError(f"{fname}: no <svg> element")
    svg_attributes = {k: _attr(svg_tag.group(0), k) for k in ("width", "height")}
    paths, attributes = [], []
    for m in re.finditer(r"<path\b[^>]*?/?>", text, re.S):
        tag = m.group(0)
        d = _attr(tag, "d")
        if d is None:
            continue
        attr = {"d": d}
        tr = _attr(tag, "transform")
        if tr is not None:
            attr["transform"] = tr
        try:
            paths.append(_path_lines(d))
        except ValueError:
            paths.append(None)  # curves: the reference skips such figures (assets.py:103-105)
        attributes.append(attr)
    return paths, attributes, svg_attributes


def write_svg(fname, figures: Sequence[Sequence[Point]], width: int, height: int, transforms=None) -> None:
    """Closed polylines -> the ``d="M x,y L x,y ..."`` form svgpathtools' ``wsvg`` writes
    (scripts/generate_maps.py:76-78).  Coordinates are SVG pixels (y down)."""
    out = [f'<?xml version="1.0" ?>\n<svg xmlns="http://www.w3.org/2000/svg" width="{width}px" height="{height}px">']
    for k, fig in enumerate(figures):
        pts = list(fig) + [fig[0]]
        d = "M " + " L ".join(f"{x!r},{y!r}" for x, y in pts)
        tr = f' transform="{transforms[k]}"' if transforms and transforms[k] else ""
        out.append(f'  <path d="{d}" fill="none" stroke="#000000"{tr}/>')
    out.append("</svg>\n")
    Path(fname).write_text("\n".join(out))


# ------------------------------------------------------------------------------------------
# reference helpers
# ------------------------------------------------------------------------------------------
def line_begin_end(line: Tuple[complex, complex], scale: float, width: int, height: int) -> List[Point]:
    """assets.py:10-25: SVG pixels (y down) -> metres (y up)."""
    a, b = line
    return [(a.real / scale, (height - a.imag) / scale), (b.real / scale, (height - b.imag) / scale)]


def sort_segments(segments: onp.ndarray, orientation: str) -> onp.ndarray:
    """assets.py:28-82: chain the segments of one closed figure head to tail, starting from the
    left-most / lowest one, so that ``[s[0] for s in result]`` walks the outline in ``orientation``."""
    if orientation not in ("counterclockwise", "clockwise"):
        raise ValueError(f"Unknown orientation {orientation}")
    seg = onp.array(segments, dtype=onp.float64, copy=True)
    n = len(seg)
    by_corner = sorted(range(n), key=lambda i: (min(seg[i, 0, 0], seg[i, 1, 0]), min(seg[i, 0, 1], seg[i, 1, 1])))
    # the two left-most candidates, each read left to right: start with the one that ends lower (:47-55)
    cand = seg[by_corner[:2]].copy()
    for c in cand:
        if c[0, 0] > c[1, 0]:
            c[0, 0], c[1, 0] = c[1, 0], c[0, 0]
    cur = by_corner[0] if (len(cand) < 2 or cand[0, 1, 1] < cand[1, 1, 1]) else by_corner[1]
    s = seg[cur].copy()
    flip = s[0, 0] > s[1, 0] or (s[0, 0] == s[1, 0] and s[0, 1] < s[1, 1])  # :57-63
    first, second = (s[1], s[0]) if flip else (s[0], s[1])
    seg[cur] = onp.stack([first, second] if orientation == "counterclockwise" else [second, first])
    order, remaining = [cur], set(range(n)) - {cur}
    while remaining:  # :72-81: nearest endpoint to the current tail comes next
        tail = seg[cur, 1]
        cur = min(remaining, key=lambda i: min(onp.linalg.norm(seg[i, j] - tail) for j in range(2)))
        if not onp.allclose(seg[cur, 0], tail):
            seg[cur] = seg[cur][::-1].copy()
        order.append(cur)
        remaining.remove(cur)
    return seg[order]


# ---- ear clipping as tripy implements it --------------------------------------------------
_EPS = math.sqrt(sys.float_info.epsilon)


def _tri_sum(x1, y1, x2, y2, x3, y3):
    return x1 * (y3 - y2) + x2 * (y1 - y3) + x3 * (y2 - y1)


def _tri_area(x1, y1, x2, y2, x3, y3):
    return abs((x1 * (y2 - y3) + x2 * (y3 - y1) + x3 * (y1 - y2)) / 2.0)


def _inside(p, a, b, c):
    area = _tri_area(a[0], a[1], b[0], b[1], c[0], c[1])
    a1 = _tri_area(p[0], p[1], b[0], b[1], c[0], c[1])
    a2 = _tri_area(p[0], p[1], a[0], a[1], c[0], c[1])
    a3 = _tri_area(p[0], p[1], a[0], a[1], b[0], b[1])
    return abs(area - (a1 + a2 + a3)) < _EPS


def _is_ear(p1, p2, p3, polygon):
    no_points = all(pn in (p1, p2, p3) or not _inside(pn, p1, p2, p3) for pn in polygon)
    convex = _tri_sum(p1[0], p1[1], p2[0], p2[1], p3[0], p3[1]) < 0
    return no_points and convex and _tri_area(p1[0], p1[1], p2[0], p2[1], p3[0], p3[1]) > 0


def earclip(polygon: Sequence[Point]) -> List[Tuple[Point, Point, Point]]:
    """Simple polygon -> triangles ``((x,y),(x,y),(x,y))``, each (prev, ear, next) in the polygon's
    (counter-clockwise) walking order; ears are clipped first-found-first."""
    poly = [(float(x), float(y)) for x, y in polygon]
    if sum((poly[(i + 1) % len(poly)][0] - poly[i][0]) * (poly[(i + 1) % len(poly)][1] + poly[i][1])
           for i in range(len(poly))) > 0:  # clockwise input: reverse
        poly.reverse()
    n = len(poly)
    ears = [poly[i] for i in range(n) if _is_ear(poly[i - 1], poly[i], poly[(i + 1) % n], poly)]
    triangles = []
    while ears and n >= 3:
        ear = ears.pop(0)
        i = poly.index(ear)
        prev_pt, next_pt = poly[i - 1], poly[(i + 1) % n]
        poly.remove(ear)
        n -= 1
        triangles.append((prev_pt, ear, next_pt))
        if n > 3:
            pi = i - 1
            groups = [(poly[pi - 1], prev_pt, next_pt), (prev_pt, next_pt, poly[(i + 1) % n])]
            for g in groups:
                p = g[1]
                if _is_ear(g[0], g[1], g[2], poly):
                    if p not in ears:
                        ears.append(p)
                elif p in ears:
                    ears.remove(p)
    return triangles


def rotate_polygon(polygon: Polygon, angle: float, center: Point) -> Polygon:
    """jax_geometry.py:264-278 (degrees, about ``center``); float32 like the reference's jnp arrays."""
    rad = onp.float32(angle / 180.0 * onp.pi)
    c, s = onp.cos(rad, dtype=onp.float32), onp.sin(rad, dtype=onp.float32)
    rot = onp.array(((c, -s), (s, c)), dtype=onp.float32)
    tr = onp.array(center, dtype=onp.float32)
    seg = onp.asarray(polygon.segments, dtype=onp.float32)
    out = ((rot @ (seg.reshape(-1, 2) - tr).T).T + tr).reshape(-1, 2, 2)
    return Polygon(segments=out.astype(onp.float32))


def get_svg_scene(fname, px_per_meter: float = 50) -> JaxScene:
    """assets.py:85-131: every closed straight-line figure -> CCW outline -> ear-clipped triangles ->
    3-segment polygons ``[(p0,p1),(p1,p2),(p2,p0)]``, concatenated polygon-major."""
    polygons = []
    paths, attributes, svg_attributes = read_svg_paths(fname)
    w = int(float(str(svg_attributes["width"]).replace("px", "")))
    h = int(float(str(svg_attributes["height"]).replace("px", "")))
    for path, attr in zip(paths, attributes):
        if path is None:
            print("Only simple line figures are currently allowed, skipping")
            continue
        if not path or path[0][0] != path[-1][1]:
            print(f"Found non-closed path {attr['d'][:40]}..., skipping")
            continue
        outline = sort_segments(
            onp.concatenate([onp.array(line_begin_end(ln, px_per_meter, w, h))[onp.newaxis] for ln in path], axis=0),
            orientation="counterclockwise",
        )
        idxs = onp.array([0, 1, 2])
        for tri in earclip([s[0] for s in outline]):
            seg = onp.zeros((3, 2, 2), dtype=outline.dtype)
            seg[:, 0] = onp.asarray(tri)
            seg[:, 1] = seg[(idxs + 1) % 3, 0]
            poly = create_polygon(seg.astype(onp.float32))
            tr = attr.get("transform", "")
            if "rotate" in tr:
                nums = [float(v) for v in re.findall(_NUM, tr)]
                angle, cx, cy = nums[0], nums[1], h - nums[2]
                poly = rotate_polygon(poly, angle, (cx / px_per_meter, cy / px_per_meter))
            polygons.append(poly)
    return create_scene(polygons)


def save_scene(fname, scene: JaxScene) -> None:
    onp.save(fname, onp.asarray(scene.segments, dtype=onp.float32))


def load_scene(fname) -> JaxScene:
    seg = onp.load(fname).astype(onp.float32)
    if seg.ndim != 3 or seg.shape[1:] != (2, 2) or seg.shape[0] % 3:
        raise ValueError(f"{fname}: expected [3T,2,2] segments, got {seg.shape}")
    return create_scene([create_polygon(seg[i:i + 3]) for i in range(0, len(seg), 3)])
