for w in "" "--workload c1_1600x1200_3step_6bit_vh --steps 5" "--workload c2_1080p_3step_8bit_v --steps 5" "--exact-triangulation"; do
timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-rowshard $w 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print(d['config']['workload'], d['config']['triangulation'][:20], round(d['roofline']['avg_launch_us'],1), round(d['roofline']['frac'],3), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
