"""cli/prepare_tokens.py -- drop-in for the reference entry point (cli/prepare_tokens.py:14-53): features jsonl ->
`{"file_name", "audio_repr"}` jsonl written to `<out_path>/<basename(data_path)>`; bad lines are skipped with a warning.

    python cli/prepare_tokens.py data_path=<features.jsonl> out_path=<dir>"""
import json
import logging
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from slamkit_b200.config import load_config, require  # noqa: E402
from slamkit_b200.tokeniser import B200UnitTokeniser  # noqa: E402

DROP = ("units", "duration", "text", "aligned_text", "split_sentence")


def process_jsonl(line: str, tokeniser):
    try:
        cur = json.loads(line)
        cur["audio_repr"] = tokeniser.stringify_representation([cur], mode="train")[0]
        for k in DROP:
            cur.pop(k, None)
        return json.dumps(cur)
    except Exception as e:  # reference behaviour: warn and skip
        logging.warning(f"Failed to process {line}. Error: {e}, skipping")
        return None


def main(argv=None):
    cfg = load_config("prepare_tokens", argv if argv is not None else sys.argv[1:])
    require(cfg, "data_path", "out_path")
    p = cfg.tokeniser.params
    tok = B200UnitTokeniser(None, dedup=p.dedup, bos_eos_token_id=p.get("bos_eos_token_id", 1), pad_token_id=p.pad_token_id,
                            num_units=p.get("num_units") or cfg.tokeniser.feature_extractor.num_units, load_fe=False)
    os.makedirs(cfg.out_path, exist_ok=True)
    out_path = f'{cfg.out_path}/{cfg.data_path.split("/")[-1]}'
    if os.path.exists(out_path):
        logging.warning(f"{out_path} already exists. Deleting it!")
        os.remove(out_path)
    with open(cfg.data_path) as f_in, open(out_path, "a+") as f_out:
        for line in f_in:
            js = process_jsonl(line, tok)
            if js:
                f_out.write(js + "\n")
    return out_path


if __name__ == "__main__":
    main()
