"""Control-plane ingest batching: `POST /v3/metric` fan-in.

The reference handler (control/endpoints.go:109-129) decodes the body into `map[string]interface{}` and calls
`bus.Publish(events.Event{events.Metric, fmt.Sprintf("%v|%v", key, value)})` once per key — N lock acquisitions and
N fan-outs for one request.  Here the whole request becomes ONE batch: the events are built on the host and handed to
`cpbus_publish` as an array, so they share a staging pass and (up to batch_cap) a single fan-out launch.  Per-subscriber
order is the order of the batch; Go's own order is the (random) map iteration order, so the reference's test compares
multisets (control/endpoints_test.go:104-145) and so do ours.

Only the event construction lives here; sockets and HTTP routing stay in the control plane (out of scope).
"""
from __future__ import annotations

import json
from decimal import Decimal

from . import events as ev

StatusOK, StatusUnprocessableEntity = 200, 422      # net/http constants used by the handler


def _go_float(x: float) -> str:
    """fmt's %v of a float64 = strconv.FormatFloat(x, 'g', -1, 64): shortest round-trip digits; %e form when the decimal
    exponent is < -4 or >= 6 (the precision used for that decision when the shortest form was asked for), so
    1000000.0 prints as 1e+06 and 123456.0 as 123456."""
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "+Inf" if x > 0 else "-Inf"
    if x == 0:
        return "-0" if str(x).startswith("-") else "0"
    sign, digits, exp = Decimal(repr(x)).as_tuple()          # repr() = shortest round-trip digits, like strconv
    digits = list(digits)
    while len(digits) > 1 and digits[-1] == 0:               # normalise: value = 0.d1d2.. * 10^dp
        digits.pop(); exp += 1
    nd = len(digits)
    dp = nd + exp
    e10 = dp - 1
    ds = "".join(map(str, digits))
    if e10 < -4 or e10 >= 6:
        mant = ds[0] + ("." + ds[1:] if nd > 1 else "")
        out = f"{mant}e{'-' if e10 < 0 else '+'}{abs(e10):02d}"
    elif dp <= 0:
        out = "0." + "0" * (-dp) + ds
    elif dp >= nd:
        out = ds + "0" * (dp - nd)
    else:
        out = ds[:dp] + "." + ds[dp:]
    return ("-" if sign else "") + out


def go_sprint_v(v) -> str:
    """fmt.Sprintf("%v", v) for the dynamic types encoding/json produces in an interface{}:
    float64, string, bool, nil, []interface{}, map[string]interface{}."""
    if v is None:
        return "<nil>"
    if v is True:
        return "true"
    if v is False:
        return "false"
    if isinstance(v, (int, float)):
        return _go_float(float(v))                           # every JSON number decodes to float64
    if isinstance(v, str):
        return v
    if isinstance(v, list):
        return "[" + " ".join(go_sprint_v(x) for x in v) + "]"
    if isinstance(v, dict):                                  # fmt prints maps as map[k:v ...]; keys sorted (Go >= 1.12; random before)
        return "map[" + " ".join(f"{k}:{go_sprint_v(v[k])}" for k in sorted(v)) + "]"
    raise TypeError(f"not a JSON value: {type(v)!r}")


def _has_inf(v) -> bool:
    if isinstance(v, float):
        return v in (float("inf"), float("-inf"))
    if isinstance(v, list):
        return any(_has_inf(x) for x in v)
    if isinstance(v, dict):
        return any(_has_inf(x) for x in v.values())
    return False


def _valid_utf8(s: str) -> str:
    """an unpaired \\uD800-style escape decodes to U+FFFD in Go (encoding/json decode.go: unquote)"""
    return s.encode("utf-16", "surrogatepass").decode("utf-16", "replace")


def _reject_constant(name):
    raise ValueError(f"invalid JSON literal {name}")         # encoding/json has no NaN / Infinity


def metric_events(body) -> list | None:
    """The events PostMetric would publish for this request body, in document order; None = the body does not decode
    into a map[string]interface{} (the handler answers 422)."""
    if isinstance(body, (bytes, bytearray)):
        # encoding/json replaces EVERY invalid byte inside a string with its own U+FFFD (utf8.DecodeRune returns RuneError
        # with width 1), where Python's errors="replace" folds a truncated multi-byte sequence into a single one:
        # b'\xe2\x82A' is two U+FFFD + 'A' in Go.  surrogateescape keeps one (lone) surrogate per bad byte, and
        # _valid_utf8 below turns each into U+FFFD; outside strings they are a syntax error, as the raw bytes are for Go.
        body = bytes(body).decode("utf-8", errors="surrogateescape")
    try:
        # parse_int=float: every JSON number decodes to float64 in Go, so `-0` stays negative zero and prints "-0"
        doc = json.loads(body, parse_constant=_reject_constant, parse_int=float)
    except (ValueError, RecursionError):
        return None
    if doc is None:                                          # `null` unmarshals into a nil map without error: nothing to publish
        return []
    if not isinstance(doc, dict):
        return None
    if any(_has_inf(v) for v in doc.values()):               # a number float64 cannot hold: Unmarshal fails in Go
        return None
    try:
        return [ev.Event(ev.Metric, _valid_utf8(f"{go_sprint_v(k)}|{go_sprint_v(v)}")) for k, v in doc.items()]
    except OverflowError:                                    # same, for integer literals beyond float64
        return None


def post_metric(bus, body):
    """Endpoints.PostMetric (control/endpoints.go:112-129): returns (None, status).  All of the request's events are
    published as one batch (`EventBus.PublishMany` -> one `cpbus_publish` call)."""
    events = metric_events(body)
    if events is None:
        return None, StatusUnprocessableEntity
    if events:
        bus.PublishMany(events)
    return None, StatusOK
