"""Service layer (N2) on CPU with a stand-in synthesizer: the HTTP route and form fields of tools/diffusion/flask_api.py,
the framed TCP protocol of tools/diffusion/tcp_api.py, and the merging of concurrent requests into one batched call."""
import http.client
import socket
import threading

import numpy as np

from fish_diffusion_b200 import service as S


class FakeSynth:
    """features [T,E] (numpy) -> wav of T*hop samples whose value encodes the item (so mix-ups are visible)."""
    hop = 4

    def __init__(self):
        self.calls = []

    def __call__(self, feats, f0s):
        self.calls.append(len(feats))
        return [np.full(f.shape[0] * self.hop, float(f[0, 0]) + 0.001 * float(p[0]), dtype=np.float32) for f, p in zip(feats, f0s)]


def frontend(audio, sr, pitch_adjust, speaker_id):
    """two segments per request; the 'features' carry the request's identity"""
    n = len(audio)
    tag = float(audio[0]) if n else 0.0
    T = max(1, n // (2 * FakeSynth.hop))
    segs = []
    for k in range(2):
        f = np.full((T, 3), tag, dtype=np.float32)
        p = np.full((T,), 100.0 * speaker_id + pitch_adjust + k, dtype=np.float32)
        segs.append((f, p, T * FakeSynth.hop))
    return segs


def _multipart(fields, wav):
    b = "XBOUNDARYX"
    body = b""
    for k, v in fields.items():
        body += f'--{b}\r\nContent-Disposition: form-data; name="{k}"\r\n\r\n{v}\r\n'.encode()
    body += f'--{b}\r\nContent-Disposition: form-data; name="sample"; filename="a.wav"\r\nContent-Type: audio/wav\r\n\r\n'.encode()
    body += wav + f"\r\n--{b}--\r\n".encode()
    return body, f"multipart/form-data; boundary={b}"


def test_http_route_batches_concurrent_requests():
    synth = FakeSynth()
    worker = S.BatchingWorker(synth, window_s=0.3)
    srv = S.make_http_server(worker, frontend, host="127.0.0.1", port=0, model_sr=8000)
    port = srv.server_address[1]
    t = threading.Thread(target=srv.serve_forever, daemon=True)
    t.start()
    results = {}

    def client(i):
        audio = np.full(800, 0.1 * (i + 1), dtype=np.float32)
        body, ctype = _multipart({"fPitchChange": "2", "sSpeakId": str(i), "sampleRate": "8000"}, S.wav_bytes(audio, 8000))
        c = http.client.HTTPConnection("127.0.0.1", port, timeout=30)
        c.request("POST", "/voiceChangeModel", body=body, headers={"Content-Type": ctype})
        r = c.getresponse()
        results[i] = (r.status, r.getheader("Content-Type"), r.read())

    th = [threading.Thread(target=client, args=(i,)) for i in range(3)]
    [x.start() for x in th]
    [x.join() for x in th]
    srv.shutdown()
    worker.close()
    for i in range(3):
        status, ctype, data = results[i]
        assert status == 200 and ctype == "audio/wav"
        out, sr = S.read_wav(data)
        assert sr == 8000 and len(out) == 800
        assert abs(out[0] - (0.1 * (i + 1) + 0.001 * (100.0 * i + 2))) < 2e-4       # this request's own segments, in order
    assert sum(synth.calls) == 6 and len(synth.calls) < 3       # 3 requests x 2 segments in fewer than 3 model calls
    # wrong route / missing file
    worker2 = S.BatchingWorker(synth)
    srv2 = S.make_http_server(worker2, frontend, host="127.0.0.1", port=0)
    threading.Thread(target=srv2.serve_forever, daemon=True).start()
    c = http.client.HTTPConnection("127.0.0.1", srv2.server_address[1], timeout=10)
    c.request("POST", "/nope", body=b"")
    assert c.getresponse().status == 404
    srv2.shutdown()
    worker2.close()


def test_tcp_frame_protocol():
    synth = FakeSynth()
    worker = S.BatchingWorker(synth, window_s=0.01)
    a, b = socket.socketpair()
    frame = 64
    th = threading.Thread(target=S.tcp_frame_loop, args=(a, worker, frontend),
                          kwargs=dict(frame_samples=frame, sr=8000, pitch_adjust=4, speaker_id=0), daemon=True)
    th.start()
    loud = np.full(frame, 0.5, dtype=np.float32)
    b.sendall(S.pack_frame(loud)[:100])           # frames may arrive in pieces (tcp_api.py:41-46)
    b.sendall(S.pack_frame(loud)[100:])
    got = b""
    while len(got) < 4 * frame:
        got += b.recv(4 * frame)
    out = S.unpack_frame(got)
    assert len(out) == frame and abs(out[0] - (0.5 + 0.001 * 4.0)) < 1e-6
    b.sendall(S.pack_frame(np.zeros(frame, dtype=np.float32)))   # silence -> zeros without a model call
    calls = len(synth.calls)
    got = b""
    while len(got) < 4 * frame:
        got += b.recv(4 * frame)
    assert not S.unpack_frame(got).any() and len(synth.calls) == calls
    b.close()
    th.join(timeout=5)
    worker.close()
