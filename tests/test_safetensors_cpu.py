"""The safetensors reader of the C ABI (mc_safetensors_*, metalchat_b200/csrc/mc_loader.cu) on the host: header parsing, entries,
shards, metadata and the malformed files the reference's parser rejects (src/safetensor.cc:83-133).  No GPU involved."""
import json
import struct

import numpy as np
import pytest

from metalchat_b200 import capi
from tests import st_files


def test_entries_dtypes_shapes_and_bytes(tmp_path):
    rng = np.random.default_rng(3)
    tensors = {
        "model.embed_tokens.weight": (rng.integers(0, 65535, size=(7, 12), dtype=np.uint16), "BF16"),
        "layers.0.attention.wq.weight": rng.integers(-128, 127, size=(4, 64), dtype=np.int8),
        "layers.0.attention.wq.scales": rng.standard_normal((4, 2)).astype(np.float32),
        "scalar": np.array(5, dtype=np.int64),
        "empty": np.zeros((0, 3), dtype=np.float32),
        "names \"quoted\" / ü": np.arange(6, dtype=np.int32).reshape(1, 2, 3),
    }
    st_files.write(tmp_path / "m.safetensors", tensors, metadata={"format": "pt", "note": "tab\there"})
    st = capi.Safetensors(tmp_path / "m.safetensors")
    assert len(st) == len(tensors)
    got = {e["name"]: e for e in st.entries()}
    assert set(got) == set(tensors)
    assert got["model.embed_tokens.weight"]["dtype"] == "BF16" and got["model.embed_tokens.weight"]["shape"] == (7, 12)
    assert got["scalar"]["shape"] == () and got["scalar"]["nbytes"] == 8
    assert got["empty"]["shape"] == (0, 3) and got["empty"]["nbytes"] == 0
    for name, t in tensors.items():
        t = t[0] if isinstance(t, tuple) else t
        np.testing.assert_array_equal(st.tensor(name), t)
    assert st.metadata("format") == "pt" and st.metadata("note") == "tab\there" and st.metadata("absent") is None
    with pytest.raises(capi.McInvalidArgument, match="no tensor named"):
        st.tensor("nope")
    st.close()


def test_directory_of_shards(tmp_path):
    a = {"x": np.arange(4, dtype=np.float32), "y": np.arange(3, dtype=np.int8)}
    b = {"z": np.arange(5, dtype=np.int32)}
    st_files.write(tmp_path / "model-00001-of-00002.safetensors", a)
    st_files.write(tmp_path / "model-00002-of-00002.safetensors", b)
    (tmp_path / "model.safetensors.index.json").write_text(json.dumps({"weight_map": {}}))
    st = capi.Safetensors(tmp_path)
    assert [e["name"] for e in st.entries()] == ["x", "y", "z"]
    np.testing.assert_array_equal(st.tensor("z"), b["z"])
    st_files.write(tmp_path / "model-00003-of-00002.safetensors", {"x": np.zeros(1, np.float32)})
    with pytest.raises(capi.McInvalidArgument, match="appears twice"):
        capi.Safetensors(tmp_path)


def _raw(tmp_path, header: bytes, data: bytes = b"", hlen=None):
    p = tmp_path / "bad.safetensors"
    p.write_bytes(struct.pack("<Q", len(header) if hlen is None else hlen) + header + data)
    return p


@pytest.mark.parametrize("header, data, hlen, what", [
    (b'{"a":{"dtype":"F32","shape":[2],"data_offsets":[0,8]}}', b"\0" * 4, None, "do not match"),          # data shorter than the offsets say
    (b'{"a":{"dtype":"F32","shape":[3],"data_offsets":[0,8]}}', b"\0" * 8, None, "do not match"),          # shape x dtype != byte range
    (b'{"a":{"dtype":"F32","shape":[2],"data_offsets":[8,0]}}', b"\0" * 8, None, "do not match"),          # begin > end
    (b'{"a":{"dtype":"Q4","shape":[2],"data_offsets":[0,2]}}', b"\0" * 2, None, "unknown dtype"),
    (b'{"a":{"dtype":"F32","shape":[2]}}', b"\0" * 8, None, "lacks dtype"),
    (b'{"a":{"dtype":"F32","shape":[2],"data_offsets":[0,8]}', b"\0" * 8, None, "malformed header"),       # unbalanced
    (b'{"a":{"dtype":"F32","shape":[-2],"data_offsets":[0,8]}}', b"\0" * 8, None, "malformed header"),
    (b'{"a":{"dtype":"F32","shape":[2],"data_offsets":[0,8]}} x', b"\0" * 8, None, "malformed header"),    # trailing bytes inside the header
    (b'{}', b"", 1 << 40, "exceeds the file"),
])
def test_malformed_files_are_rejected(tmp_path, header, data, hlen, what):
    with pytest.raises(capi.McInvalidArgument, match=what):
        capi.Safetensors(_raw(tmp_path, header, data, hlen))


def test_missing_and_short_files(tmp_path):
    with pytest.raises(capi.McInvalidArgument, match="cannot open"):
        capi.Safetensors(tmp_path / "absent.safetensors")
    p = tmp_path / "short.safetensors"
    p.write_bytes(b"\1\2\3")
    with pytest.raises(capi.McInvalidArgument, match="shorter than"):
        capi.Safetensors(p)
    (tmp_path / "d").mkdir()
    with pytest.raises(capi.McInvalidArgument, match="no \\*.safetensors file"):
        capi.Safetensors(tmp_path / "d")


def test_empty_header_and_nested_metadata(tmp_path):
    st = capi.Safetensors(_raw(tmp_path, b'{"__metadata__":{"a":"b","n":{"deep":[1,2,{"x":null}]}},"t":{"dtype":"U8","shape":[2,2],"data_offsets":[0,4],"extra":true}}   ', b"\1\2\3\4"))
    assert st.metadata("a") == "b" and st.metadata("n") is None
    np.testing.assert_array_equal(st.tensor("t"), np.array([[1, 2], [3, 4]], np.uint8))
    assert len(capi.Safetensors(_raw(tmp_path, b"{}"))) == 0


def test_random_and_mutated_headers_never_crash(tmp_path):
    """The header comes from a file somebody downloaded: whatever the bytes, the reader either loads the file or raises McInvalidArgument."""
    from hypothesis import HealthCheck, given, settings, strategies as st

    good = (b'{"__metadata__":{"format":"pt"},"a.weight":{"dtype":"BF16","shape":[2,4],"data_offsets":[0,16]},'
            b'"b":{"dtype":"F32","shape":[3],"data_offsets":[16,28]}}')
    data = bytes(range(28))
    p = tmp_path / "fuzz.safetensors"

    def attempt(header: bytes, payload: bytes, hlen=None):
        p.write_bytes(struct.pack("<Q", len(header) if hlen is None else hlen) + header + payload)
        try:
            s = capi.Safetensors(p)
        except capi.McInvalidArgument:
            return
        for e in s.entries():  # a file that loads must describe byte ranges inside itself
            s.tensor(e["name"])
        s.close()

    attempt(good, data)

    @settings(max_examples=300, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
    @given(st.integers(0, len(good) - 1), st.integers(0, 255), st.integers(0, len(good)), st.binary(max_size=40), st.integers(0, 1 << 62))
    def run(pos, byte, cut, junk, hlen):
        mutated = bytearray(good)
        mutated[pos] = byte
        attempt(bytes(mutated), data)            # one byte flipped
        attempt(good[:cut], data)                # truncated header
        attempt(good[:cut] + junk, data)         # truncated + garbage
        attempt(junk, data)                      # garbage only
        attempt(good, data[: cut % (len(data) + 1)])  # payload shorter than the offsets say
        attempt(good, data, hlen)                # header length field lies

    run()


@pytest.mark.parametrize("header, what", [
    (b'{"a\x80b":{"dtype":"U8","shape":[1],"data_offsets":[0,1]}}', "invalid UTF-8"),
    (b'{"a\x00b":{"dtype":"U8","shape":[1],"data_offsets":[0,1]}}', "control character"),
    (b'{"a\\u0000b":{"dtype":"U8","shape":[1],"data_offsets":[0,1]}}', "u0000"),
    (b'{"a\\ud800":{"dtype":"U8","shape":[1],"data_offsets":[0,1]}}', "lone surrogate"),
    (b'{"a\\udc00":{"dtype":"U8","shape":[1],"data_offsets":[0,1]}}', "lone surrogate"),
])
def test_names_that_cannot_travel_as_c_strings_are_rejected(tmp_path, header, what):
    with pytest.raises(capi.McInvalidArgument, match=what):
        capi.Safetensors(_raw(tmp_path, header, b"\1"))


def test_escaped_and_multibyte_names(tmp_path):
    st = capi.Safetensors(_raw(tmp_path, '{"t\\u00e9\\ud83d\\ude00 中":{"dtype":"U8","shape":[1],"data_offsets":[0,1]}}'.encode(), b"\7"))
    assert [e["name"] for e in st.entries()] == ["té\U0001F600 中"]
    assert int(st.tensor("té\U0001F600 中")[0]) == 7
