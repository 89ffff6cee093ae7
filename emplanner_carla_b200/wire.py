"""Packed binary form of the planner Pipe messages.

The reference pickles Python lists over a multiprocessing.Pipe once per plan, each way:
  request  test_9.py:390-392 -> :95-96   (static_obs_info [(x, y, dis)], dynamic_obs_info [(x, y, dis, speed)],
            vehicle_loc, pred_loc, vehicle_v, vehicle_a, global_frenet_path [(x, y, theta, kappa)], match_point_list)
  reply    test_9.py:220 -> :395         (path [(x, y, theta, kappa)], match_point_list, path_s, path_l)
The global route (thousands of nodes) rides along in EVERY request.  This module gives both messages a
fixed little-endian layout - float64 payloads behind a 32-byte header - so many CARLA clients can feed one GPU
planner service over any byte transport, and so requests stack into the [batch, ...] arrays of the batched API
without per-frame Python objects.  The route is optional in a request: it is identified by `route_id` and only
has to be sent when it changes.  Host-side code only (the reference's own boundary is Python); no arithmetic.

Request : header <4sHHIIIIII> = magic 'EMPQ', version, flags, route_id, n_static, n_dynamic, n_route, n_match, 0
          then float64: static[n_static,3], dynamic[n_dynamic,4], vehicle_loc[2], pred_loc[2], vehicle_v[2],
          vehicle_a[2], route[n_route,4]; then int32 match[n_match] (padded to 8 bytes)
Reply   : header <4sHHIIIIII> = magic 'EMPR', version, flags, route_id, n_path, n_sl, 0, n_match, 0
          then float64: path[n_path,4], path_s[n_sl], path_l[n_sl]; then int32 match[n_match] (padded)
"""
import struct

import numpy as np

VERSION = 1
_HDR = struct.Struct("<4sHHIIIIII")
FLAG_HAS_ROUTE = 1


class WireError(ValueError):
    pass


def _f64(a, cols):
    a = np.asarray(a, dtype="<f8")
    a = a.reshape(-1, cols) if a.size else np.zeros((0, cols), dtype="<f8")
    return a


def _ints(a):
    a = np.asarray(a, dtype="<i4").reshape(-1)
    pad = (-a.nbytes) % 8
    return a.tobytes() + b"\0" * pad


def pack_request(static_obs_info, dynamic_obs_info, vehicle_loc, pred_loc, vehicle_v, vehicle_a, global_frenet_path=None,
                 match_point_list=(), route_id=0):
    """One planning request -> bytes.  global_frenet_path=None: the service already holds route `route_id`."""
    st, dy = _f64(static_obs_info, 3), _f64(dynamic_obs_info, 4)
    route = _f64(global_frenet_path if global_frenet_path is not None else [], 4)
    match = np.asarray(match_point_list, dtype="<i4").reshape(-1)
    flags = FLAG_HAS_ROUTE if global_frenet_path is not None else 0
    head = _HDR.pack(b"EMPQ", VERSION, flags, int(route_id), len(st), len(dy), len(route), len(match), 0)
    ego = np.asarray([vehicle_loc, pred_loc, vehicle_v, vehicle_a], dtype="<f8").reshape(4, 2)
    return b"".join([head, st.tobytes(), dy.tobytes(), ego.tobytes(), route.tobytes(), _ints(match)])


def _take(buf, off, count, dtype, cols=None):
    n = count * (cols or 1)
    size = n * np.dtype(dtype).itemsize
    if off + size > len(buf):
        raise WireError("message truncated")
    a = np.frombuffer(buf, dtype=dtype, count=n, offset=off)
    return (a.reshape(count, cols) if cols else a), off + size


def unpack_request(buf):
    """bytes -> dict with the reference's field names (numpy arrays; lists of tuples via as_reference_tuple)."""
    if len(buf) < _HDR.size:
        raise WireError("message shorter than its header")
    magic, ver, flags, route_id, n_st, n_dy, n_route, n_match, _ = _HDR.unpack_from(buf, 0)
    if magic != b"EMPQ" or ver != VERSION:
        raise WireError("not an EMPQ v%d message" % VERSION)
    off = _HDR.size
    st, off = _take(buf, off, n_st, "<f8", 3)
    dy, off = _take(buf, off, n_dy, "<f8", 4)
    ego, off = _take(buf, off, 4, "<f8", 2)
    route, off = _take(buf, off, n_route, "<f8", 4)
    match, off = _take(buf, off, n_match, "<i4")
    return dict(static_obs_info=st, dynamic_obs_info=dy, vehicle_loc=ego[0], pred_loc=ego[1], vehicle_v=ego[2],
                vehicle_a=ego[3], global_frenet_path=route if flags & FLAG_HAS_ROUTE else None,
                match_point_list=match, route_id=route_id)


def as_reference_tuple(req, route=None):
    """The 8-tuple motion_planning() unpacks at test_9.py:95-96 (Python lists of tuples / floats)."""
    path = req["global_frenet_path"] if req["global_frenet_path"] is not None else route
    if path is None:
        raise WireError("request carries no route and none was supplied")
    tl = lambda a: [tuple(float(v) for v in row) for row in a]
    tt = lambda a: tuple(float(v) for v in a)
    return (tl(req["static_obs_info"]), tl(req["dynamic_obs_info"]), tt(req["vehicle_loc"]), tt(req["pred_loc"]),
            tt(req["vehicle_v"]), tt(req["vehicle_a"]), tl(path), [int(v) for v in req["match_point_list"]])


def pack_reply(path, match_point_list, path_s, path_l, route_id=0):
    """Planner reply (test_9.py:220) -> bytes."""
    p = _f64(path, 4)
    s = np.asarray(path_s, dtype="<f8").reshape(-1)
    l = np.asarray(path_l, dtype="<f8").reshape(-1)
    if len(s) != len(l):
        raise WireError("path_s and path_l differ in length")
    match = np.asarray(match_point_list, dtype="<i4").reshape(-1)
    head = _HDR.pack(b"EMPR", VERSION, 0, int(route_id), len(p), len(s), 0, len(match), 0)
    return b"".join([head, p.tobytes(), s.tobytes(), l.tobytes(), _ints(match)])


def unpack_reply(buf):
    if len(buf) < _HDR.size:
        raise WireError("message shorter than its header")
    magic, ver, _, route_id, n_path, n_sl, _, n_match, _ = _HDR.unpack_from(buf, 0)
    if magic != b"EMPR" or ver != VERSION:
        raise WireError("not an EMPR v%d message" % VERSION)
    off = _HDR.size
    p, off = _take(buf, off, n_path, "<f8", 4)
    s, off = _take(buf, off, n_sl, "<f8")
    l, off = _take(buf, off, n_sl, "<f8")
    match, off = _take(buf, off, n_match, "<i4")
    return dict(path=p, match_point_list=match, path_s=s, path_l=l, route_id=route_id)


def stack_obstacles(requests, kmax=None):
    """Static obstacles of a list of unpacked requests -> (xy [B, kmax, 2] NaN padded, count [B]): the layout the
    batched Frenet projection (HostPlanner.frenet_project) takes.  kmax = None sizes the array for the largest request
    (rounded up to an even slot count, at least 8).  A request with more obstacles than an explicit kmax is an error:
    the reference plans around EVERY obstacle it is sent (test_9.py:116-122), so none is ever dropped silently."""
    B = len(requests)
    most = max([len(r["static_obs_info"]) for r in requests], default=0)
    if kmax is None:
        kmax = max(8, most + (most & 1))
    elif most > kmax:
        raise WireError("a request carries %d static obstacles but only %d slots were asked for" % (most, kmax))
    xy = np.full((B, kmax, 2), np.nan)
    n = np.zeros(B, dtype=np.int32)
    for b, r in enumerate(requests):
        k = len(r["static_obs_info"])
        xy[b, :k] = r["static_obs_info"][:k, :2]
        n[b] = k
    return xy, n
