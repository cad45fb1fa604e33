for pr in -1 0; do echo "== SIDE_PRIO=$pr"; WACIL_SIDE_PRIO=$pr WACIL_ENGINE=layer python tools/step_parts.py 2>&1 | grep "backward, wgrad\|whole"; done
