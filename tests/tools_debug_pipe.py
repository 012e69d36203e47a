"""Diagnostic (not a test): run the pipelined kernel at a given batch and print the error word on failure."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "few-shot-transformer-tts_b200"), ROOT]
import torch  # noqa: E402
from oracle import tts_oracle as O  # noqa: E402
from tts_b200.engine import TtsEngine  # noqa: E402

B = int(os.environ.get("PB", "1"))
steps = int(os.environ.get("PN", "24"))
chunk = int(os.environ.get("PCH", "24"))
cfg = O.ModelConfig(max_generation_frames=64)
params = O.synth_params(cfg, seed=0)
params["decoder.stop_net.bias"] = torch.tensor([-1e4])
eng = TtsEngine.from_state_dict(params, cfg, "cuda:0")
batch = O.synth_batch(cfg, batch=B, text_len=int(os.environ.get("PS", "37")), n_frames=4, seed=11, ragged=B > 1)
mem = eng.encode(batch["inputs"], batch["input_lengths"], batch["input_spk_ids"], batch["input_language_vecs"])
sess = eng.new_session(B, mem.shape[1], 64, "encdec")
sess.begin(mem, batch["input_lengths"].cuda())
done = 0
while done < steps:
    sess.step(min(chunk, steps - done), impl=4)
    done += chunk
    torch.cuda.synchronize()
    c = sess.counters.cpu().tolist()
    err = sess.scratch[32 * 64:32 * 64 + 1].view(torch.int32).item()
    print("after", done, "steps: step_counter", c[0], "n_unfinished", c[1], "err", err)
    if err:
        off = 32 * 64 + 32
        dbg = sess.scratch[off:off + 2 * 16 * 160 + 64].view(torch.int64)[16 * 160:16 * 160 + 12].cpu().tolist()
        print("dbg k,base,n_tiles,f_seq,total,lane,nk,u,g,n_keys,block,fullok:", dbg)
        break
want = O.eval_batch_cached(params, cfg, batch, steps)
print("max err vs oracle", (sess.frames[:, :steps].cpu() - want["mel_pre"][:, :steps]).abs().max().item())
