"""The drop-in, end to end on the B200: the reference's UNMODIFIED client, server loop,
parser, catalog and index build (server.c parse.c db_manager.c client_context.c index.c),
linked with host/query_shim.c + libadb_b200.so in place of query.c + multimap.c
(oracle/_ref/dropin/server_b200), replays the reference's project_tests DSL files.  The
same files go through the unmodified reference pair (server_ref) on the host CPU; the two
clients must print the same bytes, test by test."""
import tempfile
import warnings

import pytest

import dsl_harness as H

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (H.ServerPair.available("ref") and H.ServerPair.available("b200")),
                                 reason="oracle/_ref/dropin not built")]
# the reference's output is not a function of its input here: uninitialised print buffer
REFERENCE_OUTPUT_UNDEFINED = {14}
TESTS = range(1, 38)


# Tests that print position / value lists in INDEX order (milestone 3).  With the index built
# on the engine (stable radix sort) the order inside a run of equal keys is ascending row order,
# with the reference's unstable quicksort it is whatever the partition steps left (SURVEY.md
# A3): those replies are compared the way the reference's own verifier compares them when exact
# match fails -- as sorted lines (infra_scripts/verify_output_standalone.sh:41-47).  With
# ADB_INDEX_BUILD=reference (the reference's index.c builds, the shim uploads) every reply is
# byte-identical.
INDEX_ORDERED = set(range(18, 32))
MODE = {"index": "engine"}


def same(t, a, b):
    if a == b:
        return True
    if MODE["index"] == "engine" and t in INDEX_ORDERED:
        return sorted(a.splitlines()) == sorted(b.splitlines())
    return False


def tie_order_garbage(t, ref_out):
    """Tests 25 and 27: the reference's own answer fails its .exp (an unclustered
    index built before a later column's clustered index keeps pre-permutation positions).  WHICH
    wrong rows it then reads depends on how its quicksort ordered equal keys of the clustered
    column, so with an engine-built (stable) index the drop-in reads other wrong rows.  Compared
    byte for byte only under ADB_INDEX_BUILD=reference."""
    return MODE["index"] == "engine" and t in INDEX_ORDERED and H.verdict(ref_out, H.exp_text(t)) == "fail"


def mismatches(ref, b200):
    return [t for t in TESTS if t not in REFERENCE_OUTPUT_UNDEFINED and not tie_order_garbage(t, ref[t])
            and not same(t, b200[t], ref[t])]


@pytest.fixture(scope="module", params=[("1", "reference"), ("1", "engine"), ("3", "engine"), ("4", "engine")],
                ids=["1gpu-reference_index", "1gpu-engine_index", "3gpus-engine_index", "4gpus-engine_index"])
def outputs(request):
    """Both pairs replay the suite.  The unmodified client reads a reply with a single recv
    (client.c:127), so a replay can come out truncated on either side for
    reasons that have nothing to do with the operators (seen once in about ten replays during
    round 1, not reproduced): a replay with mismatches is repeated, twice at most, every
    repetition is reported as a warning, and only a mismatch that persists fails the test."""
    with tempfile.TemporaryDirectory(prefix="adb_ref_") as w1, tempfile.TemporaryDirectory(prefix="adb_b200_") as w2:
        for attempt in range(3):
            ref = H.ServerPair("ref", w1).run_suite(TESTS)
            # ADB_GPUS: the unmodified server process drives that many engine contexts (one per
            # GPU; they share the device on a 1-GPU box), every column sharded over them
            MODE["index"] = request.param[1]
            b200 = H.ServerPair("b200", w2, env={"ADB_GPUS": request.param[0],
                                                 "ADB_INDEX_BUILD": request.param[1]}).run_suite(TESTS)
            bad = mismatches(ref, b200)
            if not bad:
                break
            if attempt < 2:
                warnings.warn(f"drop-in replay {attempt + 1}: tests {bad} differ "
                              f"({b200[bad[0]][:120]!r} vs {ref[bad[0]][:120]!r}); replaying")
        log = open(H.ServerPair("b200", w2).server_log, errors="replace").read()
        yield ref, b200, log


def test_drop_in_prints_what_the_reference_prints(outputs):
    ref, b200, log = outputs
    bad = mismatches(ref, b200)
    assert not bad, (bad, b200[bad[0]][:300], ref[bad[0]][:300], log[-600:])


def test_drop_in_matches_the_golden_expectations(outputs):
    ref, b200, _ = outputs
    for t in TESTS:
        vr, vb = H.verdict(ref[t], H.exp_text(t)), H.verdict(b200[t], H.exp_text(t))
        if vr != "fail":
            if MODE["index"] == "engine" and t in INDEX_ORDERED:
                assert vb != "fail", (t, vr, vb)            # tie order may turn "exact" into "sorted"
            else:
                assert vb == vr, (t, vr, vb)
    # test 14: every select is empty; the drop-in prints nothing where the reference prints
    # uninitialised bytes, which is what the .exp expects
    assert H.verdict(b200[14], H.exp_text(14)) == "exact"


def test_engine_did_the_work(outputs):
    """Every reply came from the engine: a failed operator would answer "Failed"."""
    _, b200, log = outputs
    assert not any("Failed" in o for o in b200.values())
    assert "no CPU fallback" not in log
