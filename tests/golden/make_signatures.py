"""Signatures of the reference's front-end functions ("signatures to keep verbatim"), read from the reference's
SOURCE with ``ast`` (src/util.py cannot be imported here: librosa is absent) and committed as ``ref_signatures.json``.

    python tests/golden/make_signatures.py        # needs /root/reference
"""
import ast
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = "/root/reference"

# reference file -> (mirror module, function names)
TARGETS = {
    "src/util.py": ("heart_murmur_detection_b200.util", [
        "crop_first", "random_crop", "random_mask", "random_multiply", "_butter_bandpass", "_butter_bandpass_filter",
        "_slice_data_librosa", "get_individual_segments_librosa", "get_entire_signal_librosa", "get_split_signal_librosa",
        "decide_droplast", "get_individual_cycles_librosa", "_get_lungsound_label", "_get_diagnosis_label",
        "pre_process_audio_mel_t", "split_pad_sample", "get_split_signal_fbank_pad"]),
    "src/benchmark/baseline/extract_feature.py": ("heart_murmur_detection_b200.extract_feature",
                                                  ["get_split_signal_fbank", "split_sample"]),
    "src/benchmark/baseline/vggish/vggish_input.py": ("heart_murmur_detection_b200.vggish_input", ["waveform_to_examples"]),
    "src/benchmark/baseline/hear/python/data_processing/audio_utils.py": ("heart_murmur_detection_b200.hear_input",
                                                                         ["preprocess_audio"]),
}


def signature_of(fn: ast.FunctionDef):
    a = fn.args
    pos = [x.arg for x in a.posonlyargs + a.args]
    defaults = [ast.literal_eval(d) for d in a.defaults]
    n_req = len(pos) - len(defaults)
    return {"args": pos, "defaults": {pos[n_req + i]: d for i, d in enumerate(defaults)}, "lineno": fn.lineno}


def main():
    out = {}
    for rel, (mirror, names) in TARGETS.items():
        tree = ast.parse(open(os.path.join(REFERENCE, rel)).read())
        fns = {n.name: n for n in tree.body if isinstance(n, ast.FunctionDef)}
        for name in names:
            out[f"{rel}::{name}"] = dict(signature_of(fns[name]), mirror=mirror)
    json.dump(out, open(os.path.join(HERE, "ref_signatures.json"), "w"), indent=1, sort_keys=True)
    print(len(out), "signatures")


if __name__ == "__main__":
    main()
