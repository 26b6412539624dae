nproc
for i in $(seq 1 $(nproc)); do (python -c "
import time
t=time.time()
while time.time()-t < 75: pass
" &) ; done
sleep 1
python tools/time_labeler.py --frames 5 --in-flight 4 --instances 6 --single-steps 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('loaded host, single:', d['steady_s_per_frame'], d['steady_frames_per_hour'])"
python tools/time_labeler.py --frames 5 --in-flight 4 --instances 6 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('loaded host, advance:', d['steady_s_per_frame'], d['steady_frames_per_hour'])"
sleep 40
python tools/time_labeler.py --frames 5 --in-flight 4 --instances 6 --single-steps 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('idle host, single:', d['steady_s_per_frame'], d['steady_frames_per_hour'])"
python tools/time_labeler.py --frames 5 --in-flight 4 --instances 6 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('idle host, advance:', d['steady_s_per_frame'], d['steady_frames_per_hour'])"
