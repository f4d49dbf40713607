"""The Python mirror of the Go surface (av1recon.Open / Engine.VerifyFile / Close, ProbeFile, VerifyOutput -- INTEGRATION.md): same
call shapes and error behaviour as the cgo package the daemon would link
(/root/reference/internal/daemon/daemon.go:101-115 is where its result is consumed)."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "streams")
INDEX = json.load(open(os.path.join(GOLD, "index.json")))


def test_probe_file(built):
    import av1recon
    name = "intra_10b_192x128"
    info = av1recon.ProbeFile(os.path.join(GOLD, name + ".ivf"))
    assert info.is_av1 == 1 and (info.width, info.height, info.bit_depth) == (192, 128, 10)
    with pytest.raises(RuntimeError):
        av1recon.ProbeFile(os.path.join(GOLD, "does-not-exist.mkv"))


@pytest.mark.gpu
def test_verify_file_and_verify_output(built, tmp_path):
    import av1recon
    from tools import mkvmux
    from tools.obuio import read_ivf
    name = "intra_8b_200x136"
    mkv = tmp_path / "out.av1-tmp.mkv"                      # the daemon's temp name (daemon.go:86)
    mkv.write_bytes(mkvmux.mux(read_ivf(os.path.join(GOLD, name + ".ivf")), 200, 136))
    eng = av1recon.Open(0)
    rep = eng.VerifyFile(str(mkv))
    assert rep.frames == INDEX[name]["frames"] and (rep.width, rep.height, rep.bit_depth) == (200, 136, 8)
    assert av1recon.VerifyOutput(eng, str(mkv), 199, 135).frames == rep.frames     # odd source sizes are rounded up by the transcode
    with pytest.raises(av1recon.VerifyError):
        av1recon.VerifyOutput(eng, str(mkv), 320, 240)
    bad = tmp_path / "bad.mkv"
    bad.write_bytes(mkv.read_bytes()[: mkv.stat().st_size // 2])
    with pytest.raises(av1recon.VerifyError) as ei:
        eng.VerifyFile(str(bad))
    assert ei.value.code != 0 and len(str(ei.value)) < 800                           # fits job.Reason (transcode.go:295-297)
    with pytest.raises(av1recon.VerifyError):
        eng.VerifyFile(str(tmp_path / "missing.mkv"))
    # anamorphic WebRip jobs are rescaled by the sample aspect ratio before the even-size rounding (transcode.go:93-101): an output
    # wider than the source must pass for them and still fail for ordinary jobs
    assert av1recon.VerifyOutput(eng, str(mkv), 150, 136, is_webrip_like=True).frames == rep.frames
    with pytest.raises(av1recon.VerifyError):
        av1recon.VerifyOutput(eng, str(mkv), 150, 136)
    empty = tmp_path / "empty.av1-tmp.mkv"
    empty.write_bytes(b"")
    with pytest.raises(av1recon.VerifyError):
        eng.VerifyFile(str(empty))
    eng.Close()
