#!/bin/bash
# round 2, call 52: last sanity pass on the final tree: whole GPU suite + smoke()
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2
