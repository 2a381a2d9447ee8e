"""The C-ABI boundary: libcpbus.so loads here (no GPU) and exports every symbol
include/cpbus.h declares; host-only entry points agree with the golden vectors; and
without a device the library refuses to run (no CPU fallback)."""
import ctypes as C
import json
import os
import subprocess
import re

import numpy as np

import oracle_binding as ob
from containerpilot_b200 import _native as nat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "cpbus.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cpbus_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(nat.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"libcpbus.so does not export {n}"
    assert set(names) == set(nat.SYMBOLS), set(names) ^ set(nat.SYMBOLS)
    assert nat.load().cpbus_abi_version() == 2


def test_record_layout_is_32_bytes():
    assert C.sizeof(nat.Event) == 32 and ob.EVENT_DTYPE.itemsize == 32
    assert [f[0] for f in nat.Event._fields_] == list(ob.EVENT_DTYPE.names)


def test_names_agree_with_reference_vectors():
    G = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")))
    lib = nat.load()
    for i, n in enumerate(G["code_names"]["names"]):
        assert lib.cpbus_code_name(i).decode() == n
    assert lib.cpbus_code_name(17) is None
    for name, code in G["from_string"]["accepted"].items():
        assert lib.cpbus_code_from_string(name.encode()) == code
    for name in G["from_string"]["rejected"]:
        assert lib.cpbus_code_from_string(name.encode()) == -1


def test_record_hash_spec_matches_oracle():
    lib = nat.load()
    rng = np.random.default_rng(3)
    for _ in range(200):
        e = nat.Event(int(rng.integers(0, 1 << 62)), int(rng.integers(0, 1 << 62)), int(rng.integers(0, 17)),
                      int(rng.integers(0, 1 << 32)), int(rng.integers(0, 1 << 32)), int(rng.integers(0, 4)))
        assert lib.cpbus_record_hash(C.byref(e)) == ob.lib().orc_record_hash(bytes(e))
    assert lib.cpbus_digest_multiplier() == ob.lib().orc_digest_multiplier()


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        return
    cfg = nat.Config(); cfg.n_max_subs = 8; cfg.device = -1
    h = C.c_void_p()
    assert nat.load().cpbus_create(C.byref(cfg), C.byref(h)) == nat.ENODEV
    assert not h.value
    assert b"no CPU fallback" in nat.load().cpbus_strerror(nat.ENODEV)


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under containerpilot_b200/ or include/ may reference it."""
    for base in ("containerpilot_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for fn in files:
                if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cc", ".cpp")):
                    txt = open(os.path.join(dirpath, fn), errors="ignore").read()
                    assert "oracle_binding" not in txt and "libcpbus_oracle" not in txt and "orc_" not in txt, fn


def test_create_rejects_bad_configs_before_touching_the_device():
    """Argument validation is host logic: the same status codes with or without a GPU."""
    lib = nat.load()
    def create(**kw):
        cfg = nat.Config(); cfg.n_max_subs = 8; cfg.device = -1
        for k, v in kw.items():
            setattr(cfg, k, v)
        h = C.c_void_p()
        rc = lib.cpbus_create(C.byref(cfg), C.byref(h))
        if rc == 0:
            lib.cpbus_destroy(h)
        return rc
    assert create(n_max_subs=0) == nat.EINVAL
    assert create(ring_cap=1000) == nat.EINVAL            # not a power of two
    assert create(ring_cap=32) == nat.EINVAL              # < 64
    assert create(ring_cap=1024, batch_cap=1024) == nat.EINVAL   # > ring_cap / 2
    assert create(ring_cap=1024, batch_cap=100) == nat.EINVAL    # not a multiple of 32
    assert create(ring_cap=8192, batch_cap=2048) == nat.EINVAL   # > 1024
    assert create(timers_per_sub=3) == nat.EINVAL
    assert create(store_path=9) == nat.EINVAL
    assert lib.cpbus_create(None, None) == nat.EINVAL
    assert lib.cpbus_destroy(None) == nat.EINVAL
    assert lib.cpbus_strerror(nat.EAGAIN).startswith(b"mailbox full")


C_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c")


def build_c_smoke():
    """tests/c/abi_smoke.c: the C-ABI called from plain C99 (what cgo-generated code does), -Wall -Wextra -Werror -pedantic"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(C_DIR, "abi_smoke")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(root, "include"),
                           os.path.join(C_DIR, "abi_smoke.c"), "-L", os.path.join(root, "containerpilot_b200"), "-lcpbus",
                           "-Wl,-rpath," + os.path.join(root, "containerpilot_b200"), "-o", exe])
    return exe


def test_header_is_plain_c99_and_a_c_caller_links():
    exe = build_c_smoke()
    assert os.path.exists(exe)
    import torch
    if not torch.cuda.is_available():            # without a device the program must stop at cpbus_create with ENODEV, loudly
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 1 and "no CUDA device" in r.stdout


def _mask_order(masks, active, ring_cap, block, heavy):
    lib = nat.load()
    masks = np.ascontiguousarray(masks, dtype=np.uint32)
    act = None if active is None else np.ascontiguousarray(active, dtype=np.uint8)
    out = np.zeros(len(masks), dtype=np.uint32)
    k = lib.cpbus_mask_order(masks.ctypes.data, act.ctypes.data if act is not None else None, len(masks), ring_cap, block, int(heavy), out.ctypes.data)
    return out[:k]


def test_mask_order_host_logic():
    """The walk order of the filtered fan-out (host-only): a permutation of the active subscribers, block by block of
    consecutive ids, equal masks adjacent inside a block, masks with more codes first when asked."""
    import trace as tr
    n = 20_000
    masks = tr.zipf_masks(n, 1.0, 5)
    active = (np.random.default_rng(9).random(n) < 0.9).astype(np.uint8)
    for block, heavy in ((0, False), (0, True), (0xFFFFFFFF, True), (4096, False), (4096, True), (1000, True)):
        o = _mask_order(masks, active, 1024, block, heavy)
        assert sorted(o.tolist()) == np.flatnonzero(active).tolist()                 # exactly the active ones, once each
        blk = n if block in (0, 0xFFFFFFFF) else block                                # 20,000 x 32 KiB < 16 GiB: policy = one block
        bid = o // blk
        assert (np.diff(bid.astype(np.int64)) >= 0).all()                             # blocks in id order
        for b in np.unique(bid):
            m = masks[o[bid == b]] & 0x1FFFF
            runs = np.flatnonzero(np.diff(m.astype(np.int64)) != 0).size + 1
            assert runs == np.unique(m).size                                          # every mask is ONE run inside the block
            if heavy:
                pc = np.array([bin(int(x)).count("1") for x in m])
                assert (np.diff(pc) <= 0).all()                                       # more codes first
            else:
                assert (np.diff(m.astype(np.int64)) >= 0).all()                       # plain mask order
            ids = o[bid == b]
            for mv in np.unique(m)[:20]:
                assert (np.diff(ids[m == mv].astype(np.int64)) > 0).all()             # stable: ids ascending inside a run
    # the footprint policy: 1,048,576 rings of 32 KiB = 32 GiB -> blocks of 8 GiB = 262,144 subscribers
    big = np.full(1 << 20, 0x1FFFF, dtype=np.uint32); big[1::2] = 0x2
    o = _mask_order(big, None, 1024, 0, True)
    assert len(o) == 1 << 20
    first = o[: 1 << 18]
    assert first.max() < (1 << 18) and (first[: 1 << 17] % 2 == 0).all() and (first[1 << 17:] % 2 == 1).all()
    o2 = _mask_order(big[: 1 << 19], None, 1024, 0, True)                            # 16 GiB: one global block
    assert (o2[: 1 << 18] % 2 == 0).all()
