"""Replay of the reference's planning process (test_9.py:224-227, :390-395) through the drop-in `planner` package.

Run as a script in a FRESH interpreter (tests/test_motion_planning_replay.py does that): like test_9.py's main block
it creates a multiprocessing.Pipe, FORKS the planner child before anything touches CUDA, then feeds the recorded
messages one by one and collects the payloads the child sends back.  The child executes `motion_planning(conn)` -
the reference's own source text when /root/reference exists, otherwise the restatement in motion_planning_caller.py
(proven equivalent on the CPU) - with `from planner import planning_utils, path_planning` resolving to
emplanner_carla_b200/planner (the GPU drop-in), so CUDA is first initialised inside the forked child.

usage: python tests/replay_test9.py <golden.npz> <out.npz>
"""
import multiprocessing
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "emplanner_carla_b200"))     # drop-in first on sys.path
sys.path.insert(1, HERE)

from planner import planning_utils, path_planning  # noqa: E402   (test_9.py:26, resolved by the drop-in)
import motion_planning_caller as mpc  # noqa: E402


def main(golden_path, out_path):
    assert "emplanner_carla_b200" in planning_utils.__file__, planning_utils.__file__
    g = np.load(golden_path)
    msgs = mpc.messages_from_golden(g)
    text = mpc.reference_function_text()
    source = "reference text (test_9.py)" if text is not None else "restatement (tests/motion_planning_caller.py)"
    if text is None:
        text = mpc.restated_function_text()
    motion_planning = mpc.compile_caller(text, planning_utils, path_planning)
    conn1, conn2 = multiprocessing.Pipe()
    p1 = multiprocessing.get_context("fork").Process(target=motion_planning, args=(conn2,), daemon=True)
    p1.start()
    F = len(msgs)
    out = dict(cart=np.full((F, 28, 4), np.nan), cart_n=np.zeros(F, np.int32), path_s=np.full((F, 27), np.nan),
               path_l=np.full((F, 27), np.nan), path_n=np.zeros(F, np.int32), match=np.zeros(F, np.int32),
               child_pid=np.array([p1.pid]), parent_pid=np.array([os.getpid()]))
    match = msgs[0][7]
    for f, m in enumerate(msgs):
        m = list(m)
        m[7] = match
        conn1.send(tuple(m))
        if not conn1.poll(120):
            raise SystemExit("planner child did not answer frame %d (exit code %s)" % (f, p1.exitcode))
        cart, match, ps, pl = conn1.recv()
        out["cart_n"][f] = len(cart)
        out["cart"][f, :len(cart)] = np.array(cart)
        out["path_n"][f] = len(ps)
        out["path_s"][f, :len(ps)] = ps
        out["path_l"][f, :len(pl)] = np.array(pl, dtype=np.float64).reshape(-1)
        out["match"][f] = match[0]
    p1.terminate()
    np.savez(out_path, source=source, **out)
    print("replayed %d frames through the drop-in; caller = %s" % (F, source))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
