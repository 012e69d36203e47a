"""Diagnostic (not a test): wall-clock breakdown of one synthesis job (encoder / begin / decode loop / postnet)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "few-shot-transformer-tts_b200"), ROOT]
import torch  # noqa: E402
from oracle import tts_oracle as O  # noqa: E402
from tts_b200.engine import TtsEngine  # noqa: E402

cfg = O.ModelConfig(max_generation_frames=1000)
params = O.synth_params(cfg, seed=0)
params["decoder.stop_net.bias"] = torch.tensor([-1e4])
eng = TtsEngine.from_state_dict(params, cfg, "cuda:0")
batch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in O.synth_batch(cfg, batch=32, text_len=258, n_frames=4, seed=1).items()}
sess = eng.new_session(32, 258, 1000, "encdec")


def tick():
    torch.cuda.synchronize()
    return time.perf_counter()


for rep in range(4):
    t0 = tick()
    mem = eng.encode(batch["inputs"], batch["input_lengths"], batch["input_spk_ids"], batch["input_language_vecs"])
    t1 = tick()
    sess.begin(mem, batch["input_lengths"])
    t2 = tick()
    done = 0
    while done < 1000:
        sess.step(50)
        done += 50
        left = int(sess.counters[1].item())
    t3 = tick()
    lengths = sess.lengths.clone()
    mels = sess.frames[:, :1000].contiguous()
    t4 = tick()
    aft = eng.postnet(mels, lengths, add_input=True)
    t5 = tick()
    out = eng.generate(batch, max_frames=1000, record_align="encdec", chunk=50, session=sess)
    t6 = tick()
    print("encode %.2f  begin %.2f  decode loop %.2f  copy %.2f  postnet %.2f  | sum %.2f  generate() %.2f ms" % (
        1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3), 1e3 * (t5 - t4), 1e3 * (t5 - t0), 1e3 * (t6 - t5)))
