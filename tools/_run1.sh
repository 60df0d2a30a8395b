python examples/config2_age_fit.py PLM7 256 2>&1 | tail -14
python examples/config2_age_fit.py PLM1 3 2>&1 | tail -14
python -c "import __graft_entry__ as g; g.smoke()"
