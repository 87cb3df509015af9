"""OpenAI-compatible TTS server — the counterpart of the reference's `examples/openai_server.py` (same endpoints, flags and
environment variables: POST /v1/audio/speech, GET /health, --model / --voices / --ref-audio / --ref-text / --language / --host / --port,
QWEN_TTS_MODEL / QWEN_TTS_VOICES / QWEN_TTS_REF_AUDIO / QWEN_TTS_REF_TEXT / QWEN_TTS_LANGUAGE), served by the continuous-batching
scheduler instead of a lock around a bs = 1 model.  All of it lives in `qwen3_tts_cuda_graphs_b200/server.py`; this file is the entry point
under the reference's name.

    python examples/openai_server.py --model <dir | synthetic://0.6B-Base> --ref-audio voice.wav --ref-text "..." --port 8000
    curl http://localhost:8000/v1/audio/speech -H 'Content-Type: application/json' \
         -d '{"model": "tts-1", "input": "Hello!", "voice": "alloy", "response_format": "wav"}' -o out.wav
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from qwen3_tts_cuda_graphs_b200.server import main  # noqa: E402

if __name__ == "__main__":
    main()
