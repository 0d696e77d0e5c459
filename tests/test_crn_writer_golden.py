"""The .CRN writer back-end and the DDS header against committed golden vectors generated from the unmodified reference
(tests/golden/make_crn_writer_golden.py): no reference library needed at test time, so this also runs on the GPU box.
The file must equal the reference's crn_compress output byte for byte; the port decoder (oracle/port) must decode it."""
import base64
import ctypes

import numpy as np
import pytest

import crunch2_b200 as crn
import helpers

GOLD = helpers.golden("crn_writer_golden.json")
DT = dict(endpoint_indices=np.uint16, selector_indices=np.uint16, color_endpoints=np.uint32, alpha_endpoints=np.uint32, color_selectors=np.uint32, alpha_selectors=np.uint64)


@pytest.mark.parametrize("case", GOLD["writer"], ids=[c["name"] for c in GOLD["writer"]])
def test_writer_reproduces_the_reference_file(case, port):
    lib = crn.load_library()
    p = crn.crn_params(helpers.CRN_FMT[case["name"]], case["w"], case["h"], case["nlev"], 1, case["q"], lib=lib)
    hp = crn.crn_hc_params(p, lib=lib)
    assert [hp.color_endpoint_codebook_size, hp.color_selector_codebook_size, hp.alpha_endpoint_codebook_size, hp.alpha_selector_codebook_size] == case["codebook_sizes"]
    out = {k: np.frombuffer(base64.b64decode(case[k]), dt) for k, dt in DT.items()}
    out["endpoint_indices"] = out["endpoint_indices"].reshape(-1, 4); out["selector_indices"] = out["selector_indices"].reshape(-1, 4)
    assert len(out["endpoint_indices"]) == hp.num_blocks
    want = base64.b64decode(case["file"])
    got = crn.crn_write(p, hp, out, lib=lib)
    assert got == want
    assert len(helpers.port_unpack_all(port, got)) == case["nlev"]


@pytest.mark.parametrize("case", GOLD["dds_header"], ids=["%s-%dx%d-%s-%d" % (c["fmt"], c["w"], c["h"], c["levels"], c["faces"]) for c in GOLD["dds_header"]])
def test_dds_header_is_the_reference_header(case):
    lib = crn.load_library()
    levels = case["levels"]
    if levels is None:
        levels = 1
        while (max(case["w"], case["h"]) >> levels) > 0:
            levels += 1
    buf = (ctypes.c_uint8 * 128)()
    assert lib.crn_gpu_dds_header(helpers.CRN_FMT[case["fmt"]], case["w"], case["h"], levels, case["faces"], buf) == 0
    assert bytes(buf) == base64.b64decode(case["header"])
