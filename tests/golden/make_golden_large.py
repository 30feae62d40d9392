"""Generate tests/golden/golden_large.json: BASELINE.json's configs 2-4 at their stated sizes, the digital-silence cases and the
exact sync positions, all from the UNMODIFIED reference sources built by oracle/Makefile.ref
  oracle/_ref/audiowmark   the reference CLI
  oracle/_ref/sync_dump    the reference's SyncFinder::search behind a print loop (oracle/ref_shims/sync_dump.cc)

Run here (needs /root/reference to build the binaries; about 10 minutes):   python tests/golden/make_golden_large.py
Only hashes, JSON documents and score lists are stored; the tests regenerate the inputs from seeds with the oracle
(bit exact against these hashes) on whatever box they run.
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import awm_oracle as O      # noqa: E402
import awm_testlib as T     # noqa: E402
import build_oracle         # noqa: E402

REF = build_oracle.build_reference()
DUMP = build_oracle.REF_SYNC_DUMP
assert REF and os.path.exists(REF) and os.path.exists(DUMP), "reference binaries not available"
PAYLOAD = "f0f0f0f0f0f0f0f0f0f0f0f0f0f0f0f0"


def run(exe, *args, ok_codes=(0,)):
    t = time.time()
    p = subprocess.run([exe] + [str(a) for a in args], capture_output=True, text=True)
    assert p.returncode in ok_codes, (args, p.returncode, p.stderr)
    print("  %-60s %.1f s" % (" ".join([os.path.basename(exe)] + [str(a) for a in args])[:60], time.time() - t), flush=True)
    return p


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def pcm16(path):
    x, rate, bits = O.read_wav(path)
    return O.quantize_sndfile16(x)


def sync_dump(path):
    """-> list of searches: {"mode", "n_frames", "scores": [[index, quality, "A"|"B"], ...]} in call order"""
    out = []
    for line in run(DUMP, path).stdout.splitlines():
        w = line.split()
        if w[0] == "search":
            out.append({"mode": w[1], "n_frames": int(w[2]), "scores": []})
        elif w[0] == "score":
            out[-1]["scores"].append([int(w[1]), float(w[2]), w[3]])
    return out


def get_case(tmp, wav, extra=()):
    js = os.path.join(tmp, "out.json")
    g = run(REF, "get", *extra, "--json", js, wav)
    return {"get_stdout": g.stdout, "json": json.load(open(js)), "get_args": list(extra)}


def main():
    G = {"reference": "swesterfeld/audiowmark 0.6.5 sources compiled unmodified by oracle/Makefile.ref (FFT: oracle/ref_shims/fftw_shim.cc, "
                      "resampler: oracle/ref_shims/awm_vresampler.hh); sync positions from oracle/ref_shims/sync_dump.cc"}
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as tmp:
        p = lambda name: os.path.join(tmp, name)
        # ---- digital silence: every sync quality is exactly 0, every soft bit 0/0
        O.write_wav16(p("z170.wav"), np.zeros((170 * 44100, 2), np.float32))
        G["silence170"] = dict(get_case(tmp, p("z170.wav")), seconds=170, sync=sync_dump(p("z170.wav")))
        # 60 s of watermarked noise followed by 60 s of zeros (short file: clip decoder runs as well)
        x60 = O.int16_to_float(O.quantize_sndfile16(T.noise(60.0, 2, seed=4321)))
        O.write_wav16(p("n60.wav"), x60)
        run(REF, "add", p("n60.wav"), p("n60wm.wav"), PAYLOAD)
        y60 = pcm16(p("n60wm.wav"))
        ns = np.concatenate([y60, np.zeros((60 * 44100, 2), np.int16)])
        O.write_wav16(p("ns.wav"), O.int16_to_float(ns))
        G["noise60_silence60"] = dict(get_case(tmp, p("ns.wav")), noise_seed=4321, wm_sha256=sha(y60), input_sha256=sha(ns), sync=sync_dump(p("ns.wav")))

        # ---- config 2: 1 h stereo reference noise (test-gen-noise), add + get, 3 chunks
        run(REF, "test-gen-noise", p("h.wav"), 3600, 44100)
        a = run(REF, "add", p("h.wav"), p("hwm.wav"), PAYLOAD)
        h16 = pcm16(p("hwm.wav"))
        G["hour"] = dict(get_case(tmp, p("hwm.wav")), input_sha256=sha(pcm16(p("h.wav"))), output_sha256=sha(h16), add_stderr=a.stderr,
                         payload=PAYLOAD, sync=sync_dump(p("hwm.wav")))
        c = run(REF, "cmp", p("hwm.wav"), PAYLOAD, ok_codes=(0, 1))
        G["hour"]["cmp_tail"] = c.stdout.splitlines()[-2:]
        # first 140 s of the output: what the test-clip case below and cheap tests can regenerate without the whole hour
        G["hour"]["output_head140_sha256"] = sha(h16[:140 * 44100])
        G["hour"]["output_head600_sha256"] = sha(h16[:600 * 44100])

        # ---- config 3: 30 s clip cut from the 1 h output by the reference's test-clip (seed 0)
        run(REF, "test-clip", p("hwm.wav"), p("clip.wav"), 0, 30)
        c16 = pcm16(p("clip.wav"))
        # locate the cut (start point is a keyed random position inside the first two blocks, src/audiowmark.cc:362-372)
        first = h16[:2 * 2226 * 1024 + 31 * 44100]
        key = c16[:64].tobytes()
        start = None
        hb = first.tobytes()
        pos = hb.find(key)
        while pos >= 0:
            if pos % 4 == 0 and np.array_equal(first[pos // 4: pos // 4 + len(c16)], c16):
                start = pos // 4
                break
            pos = hb.find(key, pos + 1)
        assert start is not None
        G["clip30"] = dict(get_case(tmp, p("clip.wav")), start_frame=int(start), n_frames=int(len(c16)), input_sha256=sha(c16), sync=sync_dump(p("clip.wav")))

        # ---- config 4: --detect-speed on 10 min stereo, speeds at and inside the +-10 % edges of the scan range.
        # NOTE: everything behind a resampler runs on oracle/ref_shims/awm_vresampler.hh (zita-resampler is absent).
        O.write_wav16(p("m10.wav"), O.int16_to_float(h16[:600 * 44100]))
        G["speed600"] = {"cases": []}
        for speed in (0.9, 0.9764, 1.01, 1.1):
            run(REF, "test-change-speed", p("m10.wav"), p("sp.wav"), speed)
            s16 = pcm16(p("sp.wav"))
            case = get_case(tmp, p("sp.wav"), extra=("--detect-speed",))
            case.update(speed=speed, input_sha256=sha(s16), n_frames=int(len(s16)))
            G["speed600"]["cases"].append(case)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_large.json")
    json.dump(G, open(out, "w"), indent=1, sort_keys=True)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
