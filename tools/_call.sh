# scratch command file for `gpurun -- 'bash tools/_call.sh'` (edited per experiment)
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
