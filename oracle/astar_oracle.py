"""TEST INFRASTRUCTURE (oracle): CPU restatement of the reference's A* (generate_paths.py:11-74), used only by tests/.

astar() follows the reference line by line -- binary heap of (f, (x, y)) tuples (ties broken by the position tuple, as
heapq compares them), closed set, g / h / f dictionaries, Euclidean heuristic (:62-63), step cost 1 or sqrt(2) (:66-72),
8-neighbourhood where a diagonal step needs BOTH orthogonal neighbours free (:45-59) -- so it returns the same path.
path_length() is calculate_length (test_network_and_algorithms.py:52-56), the figure the reference's results table reports.
dial_lengths() is the size-independent property the GPU kernel is checked against: the heuristic is consistent, so A*'s path
is a shortest path, and every shortest path has the same numbers of straight and diagonal steps (1 and sqrt(2) are
rationally independent), i.e. the same length up to the order of the floating-point additions.
"""
import heapq
import math

import numpy as np


def _neighbors(pos, img):  # generate_paths.py:45-59
    height, width = img.shape
    result = []
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            if dx == 0 and dy == 0:
                continue
            x, y = pos[0] + dx, pos[1] + dy
            if x < 0 or x >= height or y < 0 or y >= width:
                continue
            if dx != 0 and dy != 0:
                if img[x, pos[1]] == 0 or img[pos[0], y] == 0:
                    continue
            if img[x, y] != 0:
                result.append((x, y))
    return result


def _heuristic(a, b):  # generate_paths.py:62-63
    return math.sqrt((a[0] - b[0]) ** 2 + (a[1] - b[1]) ** 2)


def _cost(a, b):  # generate_paths.py:66-72
    return math.sqrt(2) if (abs(a[0] - b[0]) == 1 and abs(a[1] - b[1]) == 1) else 1.0


def astar(image, start, finish):  # generate_paths.py:11-42
    img = np.array(image)
    start, finish = tuple(int(v) for v in start), tuple(int(v) for v in finish)
    open_list, closed, parent = [], set(), {}
    g = {start: 0}
    f = {start: _heuristic(start, finish)}
    heapq.heappush(open_list, (f[start], start))
    while open_list:
        current = heapq.heappop(open_list)[1]
        if current == finish:
            path = []
            while current in parent:
                path.append(current)
                current = parent[current]
            path.append(start)
            path.reverse()
            return path
        closed.add(current)
        for nb in _neighbors(current, img):
            if nb in closed or img[nb] == 0:
                continue
            cand = g[current] + _cost(current, nb)
            if nb not in g or cand < g[nb]:
                parent[nb] = current
                g[nb] = cand
                f[nb] = cand + _heuristic(nb, finish)
                heapq.heappush(open_list, (f[nb], nb))
    return None


def _neighbors3d(pos, img):  # ThreeD_generate_paths.py:47-76, quirks included
    height, width, depth = img.shape
    result = []
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dz in (-1, 0, 1):
                # the reference compares the OFFSETS with 255 (`dx == 255`, `dx != 255`): never / always true.  So the zero
                # move is not skipped (astar() drops it: the current cell is in the closed set), and EVERY move, axis moves
                # included, needs the three axis-projected cells free -- a projection with a zero offset is the current cell
                x, y, z = pos[0] + dx, pos[1] + dy, pos[2] + dz
                if x < 0 or x >= height or y < 0 or y >= width or z < 0 or z >= depth:
                    continue
                if img[x, pos[1], pos[2]] == 255 or img[pos[0], y, pos[2]] == 255 or img[pos[0], pos[1], z] == 255:
                    continue
                if img[x, y, z] != 255:
                    result.append((x, y, z))
    return result


def _cost3d(a, b):  # ThreeD_generate_paths.py:83-97
    n = (abs(a[0] - b[0]) == 1) + (abs(a[1] - b[1]) == 1) + (abs(a[2] - b[2]) == 1)
    return math.sqrt(3) if n == 3 else (math.sqrt(2) if n == 2 else 1.0)


def astar3d(image, start, finish):  # ThreeD_generate_paths.py:14-44 (obstacle = 255, anything else is free)
    img = np.array(image)
    start, finish = tuple(int(v) for v in start), tuple(int(v) for v in finish)
    h3 = lambda a, b: math.sqrt((a[0] - b[0]) ** 2 + (a[1] - b[1]) ** 2 + (a[2] - b[2]) ** 2)  # noqa: E731  (:79-80)
    open_list, closed, parent = [], set(), {}
    g = {start: 0}
    heapq.heappush(open_list, (h3(start, finish), start))
    while open_list:
        current = heapq.heappop(open_list)[1]
        if current == finish:
            path = []
            while current in parent:
                path.append(current)
                current = parent[current]
            path.append(start)
            path.reverse()
            return path
        closed.add(current)
        for nb in _neighbors3d(current, img):
            if nb in closed or img[nb] == 255:
                continue
            cand = g[current] + _cost3d(current, nb)
            if nb not in g or cand < g[nb]:
                parent[nb] = current
                g[nb] = cand
                heapq.heappush(open_list, (cand + h3(nb, finish), nb))
    return None


def path_length(path):  # test_network_and_algorithms.py:52-56 (scipy's euclidean == sqrt(sum of squares) for these steps)
    length = 0  # sequential +=, as the reference (CPython's sum() of floats is compensated since 3.12: not the same bits)
    for a, b in zip(path[:-1], path[1:]):
        length += math.sqrt(sum((p - q) ** 2 for p, q in zip(a, b)))
    return float(length)


def dial_lengths(image, start):
    """Exact shortest-path step counts (straight, diagonal) from `start` to every cell: Dijkstra with a heap over the same
    move set; returns (n_straight, n_diagonal) int arrays, -1 where unreachable."""
    img = np.asarray(image) != 0
    H, W = img.shape
    INF = float("inf")
    dist = np.full((H, W), INF)
    ns = np.full((H, W), -1, np.int32)
    nd = np.full((H, W), -1, np.int32)
    s = (int(start[0]), int(start[1]))
    dist[s] = 0.0
    ns[s] = nd[s] = 0
    heap = [(0.0, s)]
    r2 = math.sqrt(2)
    while heap:
        d, u = heapq.heappop(heap)
        if d > dist[u]:
            continue
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                if dx == 0 and dy == 0:
                    continue
                x, y = u[0] + dx, u[1] + dy
                if x < 0 or x >= H or y < 0 or y >= W or not img[x, y]:
                    continue
                diag = dx != 0 and dy != 0
                if diag and (not img[x, u[1]] or not img[u[0], y]):
                    continue
                a, b = ns[u] + (0 if diag else 1), nd[u] + (1 if diag else 0)
                c = a + b * r2
                if c < dist[x, y] - 1e-9:
                    dist[x, y] = c
                    ns[x, y], nd[x, y] = a, b
                    heapq.heappush(heap, (c, (x, y)))
    return ns, nd
