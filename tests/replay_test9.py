"""Replay of the reference's planning process (test_9.py:224-227, :390-395) through the drop-in `planner` package.

Run as a script in a FRESH interpreter (tests/test_motion_planning_replay.py does that): like test_9.py's main block
it creates a multiprocessing.Pipe, FORKS the planner child before anything touches CUDA, then feeds the recorded
messages one by one and collects the payloads the child sends back.  The child executes `motion_planning(conn)` -
the reference's own source text when /root/reference exists, otherwise the restatement in motion_planning_caller.py
(proven equivalent on the CPU) - with `from planner import planning_utils, path_planning` resolving to
emplanner_carla_b200/planner (the GPU drop-in), so CUDA is first initialised inside the forked child.

Two children are forked, one after the other:
  free    the drop-in modules as they are; payloads go back through the Pipe like in production;
  forced  the drop-in modules behind motion_planning_caller.ForcingProxy: every planner call is compared with the
          reference's recorded answer to the same call and the caller continues on the reference's answer (the
          per-call report is written by the child after every frame).

usage: python tests/replay_test9.py <golden.npz> <out.npz> <forced_report.json>
"""
import json
import multiprocessing
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "emplanner_carla_b200"))     # drop-in first on sys.path
sys.path.insert(1, HERE)

from planner import planning_utils, path_planning  # noqa: E402   (test_9.py:24, resolved by the drop-in)
import motion_planning_caller as mpc  # noqa: E402


class _ReportingProxy(mpc.ForcingProxy):
    """ForcingProxy whose report is flushed to a file after the last planner call of every frame."""

    def __init__(self, module, recorded, report, cursor, path):
        super().__init__(module, recorded, report, cursor)
        self._path = path

    def __getattr__(self, name):
        call = super().__getattr__(name)
        if name != "frenet_2_x_y_theta_kappa":
            return call

        def last_call_of_frame(*a, **k):
            out = call(*a, **k)
            with open(self._path + ".tmp", "w") as f:
                json.dump({"calls_done": self._cur[0], "per_function": self._rep}, f)
            os.replace(self._path + ".tmp", self._path)
            return out
        return last_call_of_frame


def run_child(motion_planning, msgs):
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
    p1.join(10)
    return out


def main(golden_path, out_path, report_path):
    assert "emplanner_carla_b200" in planning_utils.__file__, planning_utils.__file__
    g = np.load(golden_path)
    msgs = mpc.messages_from_golden(g)
    text = mpc.reference_function_text()
    source = "reference text (test_9.py)" if text is not None else "restatement (tests/motion_planning_caller.py)"
    if text is None:
        text = mpc.restated_function_text()
    free = run_child(mpc.compile_caller(text, planning_utils, path_planning), msgs)
    recorded = json.loads(str(g["refd_calls_json"]))
    report, cursor = {}, [0]
    forced = run_child(mpc.compile_caller(text, _ReportingProxy(planning_utils, recorded, report, cursor, report_path),
                                          _ReportingProxy(path_planning, recorded, report, cursor, report_path)), msgs)
    np.savez(out_path, source=source, **free, **{"forced_" + k: v for k, v in forced.items()},
             recorded_calls=np.array([len(recorded)]))
    print("replayed %d frames twice through the drop-in; caller = %s" % (len(msgs), source))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3])
