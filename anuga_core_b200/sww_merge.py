"""Merge the per-rank SWW files of a distributed run into one global file.

Mirrors anuga/utilities/sww_merge.py: sww_merge_parallel (:26-45) and its two workers
_sww_merge_parallel_smooth (:205-540, one value per node) and _sww_merge_parallel_non_smooth
(:542-860, three values per triangle).  Triangles go to their global ids (tri_l2g) if they are full on
the rank that wrote them; node values come from the ranks whose FULL triangles touch the node, later
ranks overwriting earlier ones - as in the reference.
"""
import os

import numpy as np

from . import sww as _sww

_QUANTITIES = ("elevation", "friction", "stage", "xmomentum", "ymomentum", "xvelocity", "yvelocity", "height")


def sww_merge_parallel(domain_global_name, np_, verbose=False, delete_old=False):
    output = domain_global_name + ".sww"
    files = [domain_global_name + "_P" + str(np_) + "_" + str(v) + ".sww" for v in range(np_)]
    fid = _sww._open(files[0], "r")
    nvol = fid.dimensions["number_of_volumes"]
    npts = fid.dimensions["number_of_points"]
    fid.close()
    _merge(files, output, smooth=(3 * nvol != npts))
    if delete_old:
        for f in files:
            os.remove(f)
    return output


def _split(fid, names, n_steps):
    present = [q for q in names if q in fid.variables]
    dyn = [q for q in present if fid.variables[q].shape[0] == n_steps and len(fid.variables[q].shape) == 2]
    return [q for q in present if q not in dyn], dyn


def _merge(files, output, smooth):
    first = True
    for filename in files:
        fid = _sww._open(filename, "r")
        if first:
            times = np.array(fid.variables["time"][:])
            n_steps = len(times)
            starttime = int(fid.starttime)
            NT = int(fid.number_of_global_triangles)
            NN = int(fid.number_of_global_nodes)
            NP = NN if smooth else 3 * NT
            atts = {k: getattr(fid, k) for k in ("order", "xllcorner", "yllcorner", "zone", "false_easting",
                                                 "false_northing", "datum", "projection")}
            description = fid.description
            description = "merged:" + (description.decode() if isinstance(description, bytes) else description)
            s_q, d_q = _split(fid, _QUANTITIES, n_steps)
            s_c, d_c = _split(fid, [q + "_c" for q in _QUANTITIES], n_steps)
            g_volumes = np.zeros((NT, 3), dtype=np.int64) if smooth else np.arange(3 * NT).reshape(-1, 3)
            g_points = np.zeros((NP, 2), dtype=np.float32)
            out_s = {q: np.zeros(NP, dtype=np.float32) for q in s_q}
            out_d = {q: np.zeros((n_steps, NP), dtype=np.float32) for q in d_q}
            out_sc = {q: np.zeros(NT, dtype=np.float32) for q in s_c}
            out_dc = {q: np.zeros((n_steps, NT), dtype=np.float32) for q in d_c}
            first = False
        tri_l2g = np.array(fid.variables["tri_l2g"][:], dtype=np.int64)
        node_l2g = np.array(fid.variables["node_l2g"][:], dtype=np.int64)
        full = np.array(fid.variables["tri_full_flag"][:]) > 0
        f_ids = np.flatnonzero(full)
        f_gids = tri_l2g[f_ids]
        x = np.array(fid.variables["x"][:], dtype=np.float32)
        y = np.array(fid.variables["y"][:], dtype=np.float32)
        if smooth:
            volumes = np.array(fid.variables["volumes"][:], dtype=np.int64)
            g_volumes[f_gids] = node_l2g[volumes[f_ids]]
            g_points[node_l2g, 0] = x
            g_points[node_l2g, 1] = y
            src = np.unique(volumes[f_ids])               # nodes of this rank's full triangles
            dst = node_l2g[src]
        else:
            src = (3 * f_ids.reshape(-1, 1) + np.array([0, 1, 2])).reshape(-1)
            dst = (3 * f_gids.reshape(-1, 1) + np.array([0, 1, 2])).reshape(-1)
            g_points[dst, 0] = x[src]
            g_points[dst, 1] = y[src]
        for q in s_q:
            out_s[q][dst] = np.array(fid.variables[q][:], dtype=np.float32)[src]
        for q in d_q:
            out_d[q][:, dst] = np.array(fid.variables[q][:], dtype=np.float32)[:, src]
        for q in s_c:
            out_sc[q][f_gids] = np.array(fid.variables[q][:], dtype=np.float32)[f_ids]
        for q in d_c:
            out_dc[q][:, f_gids] = np.array(fid.variables[q][:], dtype=np.float32)[:, f_ids]
        fid.close()

    fido = _sww._open(output, "w")
    # the reference writes both kinds of merged file with the default header (smoothing 'Yes') and, for
    # uniquely stored vertices, 3*NT as the point count (sww_merge.py:464-470, 714-718)
    _sww.write_header(fido, starttime, NT, NN if smooth else 3 * NT, True, 1, s_q, d_q, s_c, d_c,
                      description=description)
    _sww.write_georeference(fido, None)
    for k, v in atts.items():
        setattr(fido, k, v)
    fido.variables["x"][:] = g_points[:, 0]
    fido.variables["y"][:] = g_points[:, 1]
    fido.variables["volumes"][:] = g_volumes.astype(np.int32)
    for q in s_q:
        fido.variables[q][:] = out_s[q]
        fido.variables[q + _sww.RANGE][0] = np.min(out_s[q])
        fido.variables[q + _sww.RANGE][1] = np.max(out_s[q])
    for q in s_c:
        fido.variables[q][:] = out_sc[q]
    for i in range(n_steps):
        fido.variables["time"][i] = times[i]
    for q in d_q:
        for i in range(n_steps):
            fido.variables[q][i] = out_d[q][i]
        rng = fido.variables[q + _sww.RANGE]
        lo, hi = np.min(out_d[q]), np.max(out_d[q])
        if lo < rng[0]:
            rng[0] = lo
        if hi > rng[1]:
            rng[1] = hi
    for q in d_c:
        for i in range(n_steps):
            fido.variables[q][i] = out_dc[q][i]
    fido.close()
