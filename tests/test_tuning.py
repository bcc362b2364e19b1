"""Tuning-file tools (CPU): table construction, text format, and agreement
between the Python writer and the native parser (CommTuning)."""

import json
from pathlib import Path

import pytest

from faabric_b200.parallel import autotune

ROOT = Path(__file__).resolve().parent.parent


def test_table_from_rows_picks_fastest_and_merges():
    rows = [
        {"bytes": 1024, "ll_us": 5.0, "oneshot_us": 7.0, "twoshot_us": 11.0},
        {"bytes": 4096, "ll_us": 5.5, "oneshot_us": 7.0, "twoshot_us": 11.0},
        {"bytes": 65536, "ll_us": 30.0, "oneshot_us": 9.0, "twoshot_us": 12.0, "nvls_us": 13.0},
        {"bytes": 1 << 20, "oneshot_us": 40.0, "twoshot_us": 20.0, "nvls_us": 15.0},
        {"bytes": 1 << 24, "twoshot_us": 90.0, "nvls_us": 60.0, "auto_us": 1.0, "ll": "error: too large"},
    ]
    table = autotune.table_from_rows(rows)
    assert table == [(4096, "ll"), (65536, "oneshot"), (autotune.U64_MAX, "nvls")]


def test_table_hysteresis_keeps_incumbent_on_noise():
    rows = [
        {"bytes": 1024, "oneshot_us": 7.0, "twoshot_us": 9.0},
        {"bytes": 4096, "oneshot_us": 8.0, "twoshot_us": 7.9},  # within 3 %
        {"bytes": 16384, "oneshot_us": 12.0, "twoshot_us": 9.0},
    ]
    assert autotune.table_from_rows(rows) == [(4096, "oneshot"), (autotune.U64_MAX, "twoshot")]
    assert autotune.table_from_rows(rows, hysteresis=0.0)[0] == (1024, "oneshot")
    assert autotune.table_from_rows([]) == []


def test_format_parse_round_trip_and_native_agreement():
    table = [(1 << 20, "twoshot"), (4096, "ll"), (autotune.U64_MAX, "nvls")]
    settings = {"tmaMinBytes": 262144, "threads": 256}
    text = autotune.format_tuning(table, settings, comment="unit test\nsecond line")
    got_table, got_settings = autotune.parse_tuning(text)
    assert got_table == sorted(table)
    assert got_settings == settings
    # the C++ parser accepts it and re-serialises to the same directives
    native = autotune.native_normalise(text)
    assert autotune.parse_tuning(native) == (sorted(table), settings)


@pytest.mark.parametrize(
    "bad",
    ["allreduce 4096 warp9\n", "allreduce many ll\n", "set nope 3\n", "hello\n", "set threads\n"],
)
def test_both_parsers_reject_malformed(bad):
    with pytest.raises(ValueError):
        autotune.parse_tuning(bad)
    with pytest.raises(ValueError, match="line 1"):
        autotune.native_normalise(bad)


def test_format_rejects_unknown_names():
    with pytest.raises(ValueError):
        autotune.format_tuning([(1, "ring")])
    with pytest.raises(ValueError):
        autotune.format_tuning([], {"bogus": 1})


def test_cli_converts_measured_json(tmp_path):
    src = ROOT / "profiles" / "tuning_N8.json"
    out = tmp_path / "t.txt"
    assert autotune.main(["--from-json", str(src), "--set", "nvlsScalarMinBytes=33554432", "--out", str(out)]) == 0
    table, settings = autotune.parse_tuning(out.read_text())
    measured = json.loads(src.read_text())["allreduce"]
    assert [a for _, a in table] == [e["algo"] for e in measured]
    assert table[-1][0] == autotune.U64_MAX
    assert settings == {"nvlsScalarMinBytes": 33554432}


@pytest.mark.parametrize("n", [2, 4, 8])
def test_committed_tables_pick_within_noise_of_the_best_algorithm(n):
    """The table measured on B200s (profiles/tuning_N*.json) never answers with
    an algorithm that the same sweep measured more than 15 % slower than the
    best one at that size - the AUTO policy may not lose to a fixed algorithm."""
    import json
    from pathlib import Path

    from faabric_b200.parallel import autotune

    path = Path(__file__).resolve().parent.parent / "profiles" / f"tuning_N{n}.json"
    if not path.exists():
        pytest.skip(f"no measured table for {n} GPUs")
    doc = json.loads(path.read_text())
    rows = doc["rows"] if "rows" in doc else doc.get("sweep", [])
    if not rows:
        pytest.skip("table file carries no sweep rows")
    table = autotune.table_from_rows(rows)
    assert table and table[-1][0] == autotune.U64_MAX
    for r in rows:
        times = autotune.row_times(r)
        if not times:
            continue
        pick = autotune.pick_for(table, int(r["bytes"]))
        assert pick in times, (r["bytes"], pick)
        best = min(times.values())
        assert times[pick] <= best * 1.15 + 0.5, (n, r["bytes"], pick, times)
