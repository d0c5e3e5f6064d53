"""`python src/inference.py in.wav out.wav [...]` -- same CLI as the reference entry point."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from ml_audio_restoration_b200.inference import restore_audio, main  # noqa: E402,F401

if __name__ == "__main__":
    main()
