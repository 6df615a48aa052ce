"""Stand-ins for the three TensorFlow objects main.py creates around the PPO graphs
(/root/reference/main.py:31-32,46): tf.Session, tf.train.Coordinator, tf.summary.FileWriter.

A "graph tensor" of the reference becomes a Fetch handle (owner PPO + name); Session.run groups
the requested handles by owner and lets each PPO evaluate them on the engine."""
from collections import OrderedDict


class Placeholder:
    """tf.placeholder stand-in: a key for feed_dict (PPO.py:17-18,26-28)."""
    def __init__(self, name, shape=None):
        self.name, self.shape = name, shape

    def __repr__(self):
        return f"<Placeholder {self.name}>"


class Fetch:
    """A fetchable graph node of one PPO object (PPO.py:21-23,34-47,53)."""
    def __init__(self, owner, kind):
        self.owner, self.kind = owner, kind

    def __repr__(self):
        return f"<Fetch {self.owner.scope}/{self.kind}>"


class Session:
    """tf.Session stand-in (main.py:31).  run(fetches, feed_dict) mirrors sess.run."""
    def run(self, fetches, feed_dict=None):
        feed_dict = feed_dict or {}
        single = not isinstance(fetches, (list, tuple))
        flist = [fetches] if single else list(fetches)
        flat = []
        for f in flist:                      # sync_pis / assign lists are lists of ops (PPO.py:47)
            flat.extend(f if isinstance(f, (list, tuple)) else [f])
        groups = OrderedDict()
        for f in flat:
            groups.setdefault(id(f.owner), (f.owner, []))[1].append(f.kind)
        results = {}
        for owner, kinds in groups.values():
            vals = owner._run(kinds, feed_dict)
            for k in kinds:
                results[(id(owner), k)] = vals.get(k)
        out = []
        for f in flist:
            if isinstance(f, (list, tuple)):
                out.append([results[(id(x.owner), x.kind)] for x in f])
            else:
                out.append(results[(id(f.owner), f.kind)])
        return out[0] if single else out


class Coordinator:
    """tf.train.Coordinator as the reference uses it: a stop flag (main.py:32,62; Chief.py:20,87)."""
    def __init__(self):
        self._stop = False

    def should_stop(self):
        return self._stop

    def request_stop(self):
        self._stop = True

    def join(self, threads=None):
        for t in threads or []:
            if hasattr(t, "join"):
                t.join()


class FileWriter:
    """tf.summary.FileWriter stand-in (main.py:46): keeps the scalars of PPO.summary_op in memory
    and appends them to <path>/dppo_scalars.csv when the directory is writable."""
    def __init__(self, path, graph=None):
        self.path, self.rows = path, []

    def add_summary(self, summary, step):
        self.rows.append((step, dict(summary)))
        try:
            import os
            os.makedirs(self.path, exist_ok=True)
            with open(os.path.join(self.path, "dppo_scalars.csv"), "a") as f:
                f.write(",".join([str(step)] + [f"{k}={v:.8g}" for k, v in summary.items()]) + "\n")
        except OSError:
            pass
