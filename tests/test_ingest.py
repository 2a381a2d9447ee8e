"""SURVEY §8f N4 — /v3/metric fan-in as one batch (containerpilot_b200/ingest.py).  CPU part: the event construction
(Go's fmt "%v" on JSON-decoded values, status codes) against the reference's own test vectors
(control/endpoints_test.go:104-145) and the documented behaviour of fmt / encoding/json."""
import json
import os
import struct
import subprocess
from collections import Counter

import numpy as np
import pytest

from containerpilot_b200 import events as ev
from containerpilot_b200 import ingest


class FakeBus:
    """records what the handler hands to the bus"""
    def __init__(self):
        self.batches = []

    def PublishMany(self, events):
        self.batches.append(list(events))


@pytest.mark.parametrize("body,status,expected", [
    ("{{\n", 422, {}),                                                          # endpoints_test.go:125-130
    ('{"mymetric": 1.0}', 200, {(ev.Metric, "mymetric|1"): 1}),                 # :131-136
    ('{"mymetric": 1.5, "myothermetric": 2}', 200,                              # :137-144
     {(ev.Metric, "mymetric|1.5"): 1, (ev.Metric, "myothermetric|2"): 1}),
])
def test_post_metric_reference_vectors(body, status, expected):
    bus = FakeBus()
    resp, got_status = ingest.post_metric(bus, body)
    assert resp is None and got_status == status
    got = Counter((e.Code, e.Source) for b in bus.batches for e in b)
    assert dict(got) == expected
    assert len(bus.batches) == (1 if expected else 0)          # one batch per request, not one publish per key


@pytest.mark.parametrize("value,text", [
    (1.0, "1"), (2, "2"), (1.5, "1.5"), (-3.25, "-3.25"), (0, "0"), (100000.0, "100000"), (123456.0, "123456"),
    (1000000.0, "1e+06"), (1234567.0, "1.234567e+06"), (123456789.0, "1.23456789e+08"), (1e21, "1e+21"),
    (0.0001, "0.0001"), (0.00001, "1e-05"), (0.000012345, "1.2345e-05"), (0.1, "0.1"), (2.5e-7, "2.5e-07"),
    (1e100, "1e+100"), (4611686018427387904, "4.611686018427388e+18"),
    ("up", "up"), (True, "true"), (False, "false"), (None, "<nil>"),
    ([1, 2.5, "x"], "[1 2.5 x]"), ([], "[]"), ({"b": 1, "a": [True]}, "map[a:[true] b:1]"),
])
def test_go_percent_v_of_json_values(value, text):
    assert ingest.go_sprint_v(value) == text


def test_bodies_the_go_decoder_rejects_or_ignores():
    bus = FakeBus()
    for body in ("", "[1, 2]", "3", '"x"', '{"a": NaN}', '{"a": Infinity}', '{"a": 1e999}', b"\xff\xfe", '{"a": 1,}'):
        assert ingest.post_metric(bus, body) == (None, 422), body
    assert ingest.post_metric(bus, "null") == (None, 200)       # nil map: the loop runs zero times
    assert ingest.post_metric(bus, "{}") == (None, 200)
    assert bus.batches == []
    assert ingest.post_metric(bus, b'{"a": 1, "a": 2, "b": "x|y"}') == (None, 200)      # duplicate key: last one wins
    assert [(e.Code, e.Source) for e in bus.batches[0]] == [(ev.Metric, "a|2"), (ev.Metric, "b|x|y")]


def test_invalid_utf8_and_lone_surrogates_become_replacement_characters():
    bus = FakeBus()
    assert ingest.post_metric(bus, b'{"a\xff": "x\xfe"}') == (None, 200)
    assert ingest.post_metric(bus, '{"s": "\\ud800z", "ok": "\\ud83d\\ude00"}') == (None, 200)
    got = [e.Source for b in bus.batches for e in b]
    assert got == ["a\ufffd|x\ufffd", "s|\ufffdz", "ok|\U0001f600"]
    for src in got:
        src.encode("utf-8")                                  # internable


HOST_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "containerpilot_b200", "csrc", "host")


def _cpp_binary():
    exe = os.path.join(HOST_DIR, "ingest_test")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", HOST_DIR, "ingest_test"])
    return exe


def test_cpp_mirror_self_test():
    """containerpilot_b200/csrc/host/ingest.hpp against the same reference vectors (no bus, no GPU)"""
    out = subprocess.run([_cpp_binary()], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and out.stdout.strip().endswith("PASS"), out.stdout + out.stderr


def test_cpp_and_python_mirrors_agree_on_random_documents():
    """two independent restatements of json.Unmarshal + fmt %v must produce the same events for the same bodies"""
    rng = np.random.default_rng(0xC0DEB2A5)

    def rand_float():
        kind = rng.integers(0, 6)
        if kind == 0:
            return float(rng.integers(-10**6, 10**7))
        if kind == 1:
            return float(rng.integers(0, 10**6)) / float(10 ** int(rng.integers(0, 9)))
        if kind == 2:                                           # any finite double, by bit pattern
            while True:
                x = struct.unpack("<d", struct.pack("<Q", int(rng.integers(0, 2**63, dtype=np.uint64)) | (int(rng.integers(0, 2)) << 63)))[0]
                if x == x and abs(x) != float("inf"):
                    return x
        if kind == 3:
            return float(10.0 ** int(rng.integers(-30, 30))) * float(rng.integers(1, 10))
        if kind == 4:
            return float(rng.random())
        return float(rng.integers(0, 100)) / 4

    def rand_value(depth=0):
        kind = rng.integers(0, 9 if depth < 2 else 6)
        if kind <= 2:
            return rand_float()
        if kind == 3:
            return ["up", "x|y", "", "h\u00e9llo \u20ac", "tab\there", "\U0001f600"][int(rng.integers(0, 6))]
        if kind == 4:
            return [True, False, None][int(rng.integers(0, 3))]
        if kind == 5:
            return int(rng.integers(-5, 5))
        if kind <= 7:
            return [rand_value(depth + 1) for _ in range(int(rng.integers(0, 4)))]
        return {f"k{int(rng.integers(0, 5))}": rand_value(depth + 1) for _ in range(int(rng.integers(0, 4)))}

    bodies = []
    for _ in range(400):
        doc = {f"metric{int(rng.integers(0, 12))}": rand_value() for _ in range(int(rng.integers(0, 8)))}
        bodies.append(json.dumps(doc, ensure_ascii=bool(rng.integers(0, 2))))
    bodies += ["{{", "", "null", "[1]", '{"a": 1e999}', '{"a": 1e-999}', '{"a": -0.0}', '{"a": 1E5, "b": 1e6, "c": 123456.7e1}',
               '{"a": "\\ud800z"}', '{"dup": 1, "dup": [2]}', '{"a":1,}', '{"n": 12345678901234567890}']
    assert all("\n" not in b for b in bodies)
    out = subprocess.run([_cpp_binary(), "--events"], input="\n".join(bodies) + "\n", capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.split("\n")[:-1]
    assert len(lines) == len(bodies)
    n_events = 0
    for body, line in zip(bodies, lines):
        evs = ingest.metric_events(body)
        want = "422" if evs is None else "\t".join(["200"] + [e.Source for e in evs])
        assert line == want, body
        n_events += len(evs or [])
    assert n_events > 1000


def test_mirrors_agree_on_malformed_bytes_and_negative_zero():
    """Edge inputs where a careless restatement drifts from encoding/json: every invalid byte inside a string becomes its
    OWN U+FFFD (Go's utf8.DecodeRune returns width 1), and `-0` is float64 negative zero, printed "-0" by %v."""
    bodies = [b'{"k":"\xe2\x82A"}', b'{"k":"\xed\xa0\x80"}', b'{"k":"\xf0\x9f\x98"}', b'{"k\xc0\xaf":"\xff\xfe\xfd"}',
              b'{"a": -0}', b'{"a": -0, "b": [-0, -0.0, 0]}', b'{"a": -0e5}', b'{"a": 100000, "b": 1000000, "c": -1000000}']
    want_first = ["k|��A", "k|���", "k|���", "k��|���", "a|-0"]
    for body, want in zip(bodies, want_first):
        assert ingest.metric_events(body)[0].Source == want, body
    out = subprocess.run([_cpp_binary(), "--events"], input=b"\n".join(bodies) + b"\n", capture_output=True, timeout=60)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.decode("utf-8").split("\n")[:-1]
    assert len(lines) == len(bodies)
    for body, line in zip(bodies, lines):
        evs = ingest.metric_events(body)
        assert line == "\t".join(["200"] + [e.Source for e in evs]), body


def test_cpp_decoder_handles_a_body_with_1e5_keys_in_linear_time():
    body = ("{" + ",".join(f'"k{i}": {i}' for i in range(100_000)) + ',"k7": "again"}').encode()
    import time
    t0 = time.perf_counter()
    out = subprocess.run([_cpp_binary(), "--events"], input=body + b"\n", capture_output=True, timeout=60)
    assert out.returncode == 0 and time.perf_counter() - t0 < 10
    fields = out.stdout.decode().rstrip("\n").split("\t")
    assert fields[0] == "200" and len(fields) == 100_001 and fields[8] == "k7|again"     # last value wins, first position stays
