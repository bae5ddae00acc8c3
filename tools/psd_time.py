import sys
sys.path.insert(0, '.')
sys.argv = ['x', 'none']
import importlib.util, torch
spec = importlib.util.spec_from_file_location('qb', 'tools/quick_bench.py'); qb = importlib.util.module_from_spec(spec)
src = open('tools/quick_bench.py').read().split("if __name__ == '__main__':")[0]
exec(compile(src, 'qb', 'exec'), qb.__dict__)
qb.time_psd(65536, 'f64', 4096); qb.time_psd(65536, 'f32', 4096); qb.time_psd(32768, 'f64', 8192); qb.time_psd(16384, 'f32', 8192)
