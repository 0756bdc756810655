set -x
echo "== direct"; python tools/chain_phases.py 1
echo "== mask"; python tools/chain_phases.py 2
