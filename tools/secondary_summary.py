import json,sys
l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
s=l['secondary']
print('value',round(l['value']), 'single',s['single_pair_640x480']['optimize_ms_device'],'ceres',s['ceres_config_640x480']['optimize_ms_device'],'vo',s['vo_sequence_640x480']['optimize_ms_device'],'8k',s['single_pair_7680x4320']['optimize_ms_device'],'waves',round(s['batch_other_solvers_640x480']['photometric_plus_depth']['pairs_per_s']),round(s['batch_other_solvers_640x480']['ceres_mode']['pairs_per_s']))
