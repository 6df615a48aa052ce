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


# ---- TensorBoard event files, written by hand (no TensorFlow here) -----------------------------------------
# An event file is a TFRecord stream: [u64 length][u32 masked crc32c(length)][payload][u32 masked crc32c(payload)],
# payload = serialized tensorflow.Event { double wall_time = 1; int64 step = 2; string file_version = 3 |
# Summary summary = 5 { repeated Value value = 1 { string tag = 1; float simple_value = 2; } } }.
def _crc32c_table():
    poly, table = 0x82F63B78, []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ poly if c & 1 else c >> 1
        table.append(c)
    return table


_CRC_TABLE = _crc32c_table()


def _masked_crc32c(data: bytes) -> int:
    c = 0xFFFFFFFF
    for b in data:
        c = _CRC_TABLE[(c ^ b) & 0xFF] ^ (c >> 8)
    c ^= 0xFFFFFFFF
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def _varint(n: int) -> bytes:
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        out.append(b | (0x80 if n else 0))
        if not n:
            return bytes(out)


def _field_bytes(num: int, payload: bytes) -> bytes:
    return _varint((num << 3) | 2) + _varint(len(payload)) + payload


def encode_scalar_event(wall_time: float, step: int, scalars: dict) -> bytes:
    import struct
    values = b""
    for tag, val in scalars.items():
        value = _field_bytes(1, str(tag).encode("utf-8")) + _varint((2 << 3) | 5) + struct.pack("<f", float(val))
        values += _field_bytes(1, value)
    return (_varint((1 << 3) | 1) + struct.pack("<d", wall_time) + _varint((2 << 3) | 0) + _varint(int(step)) +
            _field_bytes(5, values))


def encode_record(payload: bytes) -> bytes:
    import struct
    head = struct.pack("<Q", len(payload))
    return head + struct.pack("<I", _masked_crc32c(head)) + payload + struct.pack("<I", _masked_crc32c(payload))


class FileWriter:
    """tf.summary.FileWriter (main.py:46): writes a real TensorBoard event file
    <path>/events.out.tfevents.<time>.<host>.dppo_b200 holding the scalars of PPO.summary_op
    (entropyLoss / policyLoss / valueLoss / total_loss, PPO.py:41-45); `tensorboard --logdir <path>` reads it.
    The same rows are kept in memory (`rows`) and, as before, appended to <path>/dppo_scalars.csv."""
    def __init__(self, path, graph=None):
        import os, socket, struct, time
        self.path, self.rows, self._f = path, [], None
        try:
            os.makedirs(path, exist_ok=True)
            name = "events.out.tfevents.%010d.%s.dppo_b200" % (int(time.time()), socket.gethostname())
            self.event_path = os.path.join(path, name)
            self._f = open(self.event_path, "ab")
            first = (_varint((1 << 3) | 1) + struct.pack("<d", time.time()) + _field_bytes(3, b"brain.Event:2"))
            self._f.write(encode_record(first))
            self._f.flush()
        except OSError:
            self._f = None

    def add_summary(self, summary, step):
        import os, time
        self.rows.append((step, dict(summary)))
        if self._f is not None:
            self._f.write(encode_record(encode_scalar_event(time.time(), step, summary)))
            self._f.flush()
        try:
            with open(os.path.join(self.path, "dppo_scalars.csv"), "a") as f:
                f.write(",".join([str(step)] + [f"{k}={v:.8g}" for k, v in summary.items()]) + "\n")
        except OSError:
            pass

    def flush(self):
        if self._f is not None:
            self._f.flush()

    def close(self):
        if self._f is not None:
            self._f.close()
            self._f = None
