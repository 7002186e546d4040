cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests -q -m gpu -x -k "out_of_range or small or cli or inpaint" 2>&1 | tail -5
