"""SURVEY §8f N4 — /v3/metric fan-in as one batch (containerpilot_b200/ingest.py).  CPU part: the event construction
(Go's fmt "%v" on JSON-decoded values, status codes) against the reference's own test vectors
(control/endpoints_test.go:104-145) and the documented behaviour of fmt / encoding/json."""
from collections import Counter

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
