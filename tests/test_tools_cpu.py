"""The evidence tools that run without a GPU keep working on the committed files."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_launch_summary_reads_the_committed_launch_list():
    out = subprocess.run(
        [sys.executable, os.path.join(ROOT, 'tools', 'launch_summary.py'),
         os.path.join(ROOT, 'profiles', 'r1_launches_v18_step.csv')],
        capture_output=True, text=True, check=True).stdout
    lines = out.splitlines()
    assert lines[0].startswith('1472 launches, 6 steps')
    assert 'step -2: 223 launches' in lines[1]
    rows = [l for l in lines if l.startswith('| `conv_gemm_tc_kernel')]
    assert rows and 'conv_gemm_tc_kernel<256, 6, 1, 0>' in rows[0]      # the largest share
    share = sum(float(l.split('|')[4].strip().rstrip(' %')) for l in lines if l.startswith('| `'))
    assert abs(share - 100.) < 0.5


def test_bench_reference_arm_other_ranks_exit_quietly():
    """bench.py --impl reference under torchrun: rank 0 alone measures; the other ranks
    print nothing and exit 0."""
    env = dict(os.environ, OMP_NUM_THREADS='1', RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                        '--gpus', '2', '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ''


def test_bench_parses_the_committed_ncu_pages():
    """roofline.traffic and split.hbm_bound_launches.l2_to_sm come from committed ncu raw
    pages at run time: the newest conv page gives the DRAM bytes of one res5 3x3 launch, the
    conv1x1 page the L2 -> SM bytes per L2 cycle of the res5 conv3 launch."""
    sys.path.insert(0, ROOT)
    import bench
    traffic, src = bench.ncu_traffic('r*_ncu_conv_*raw.csv', 'conv_gemm_tc_kernel')
    assert src.endswith('_raw.csv') and 'conv' in src
    assert 150e6 < traffic < 260e6          # 105.5 + 9.4 + 102.8 MB algorithmic
    l2 = bench.ncu_l2_delivery('r*_ncu_conv1x1_raw.csv', 'conv_gemm_tc_kernel')
    assert l2 is not None and l2['source'].endswith('conv1x1_raw.csv')
    assert 5000 < l2['bytes_per_l2_cycle'] < 8000 and 0.8 < l2['frac'] < 1.3
    assert bench.ncu_traffic('no_such_*.csv', 'x') == (None, None)
    assert bench.ncu_l2_delivery('no_such_*.csv', 'x') is None
