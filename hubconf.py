"""torch.hub entry point with the reference's signature (reference hubconf.py:26-56).

    model, (get_speech_timestamps, save_audio, read_audio, VADIterator, collect_chunks) = silero_vad()

`onnx`, `force_onnx_cpu` and `opset_version` choose between model files / runtimes in the reference; every variant
here is the CUDA engine (there is no CPU path), so they only keep the reference's argument checking: an opset other
than 15 / 16 with onnx=True raises, and opset 15 restricts the model to 16 kHz like silero_vad_16k_op15.onnx.
"""
dependencies = ["torch"]

from silero_vad_b200 import (VADIterator, collect_chunks, get_speech_timestamps, load_silero_vad, read_audio,  # noqa: E402
                             save_audio)


def silero_vad(onnx=False, force_onnx_cpu=False, opset_version=16):
    """Silero voice activity detector on a B200: returns (model, utils) exactly as the reference hub entry does."""
    del force_onnx_cpu   # no ONNX runtime and no CPU execution provider in this implementation
    model = load_silero_vad(onnx=onnx, opset_version=opset_version)
    utils = (get_speech_timestamps, save_audio, read_audio, VADIterator, collect_chunks)
    return model, utils
