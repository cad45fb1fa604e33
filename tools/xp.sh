WACIL_ENGINE=layer python tools/step_parts.py 2>&1 | grep "student\|backward\|whole"
