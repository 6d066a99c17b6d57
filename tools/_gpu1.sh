cd /root/repo
export KB_WATCHDOG_S=60
for P in 16,8,0 24,16,8 28,16,255 16,8,255 28,16,8 32,24,16 8,8,8 28,28,28; do
  echo "== plan $P"
  KB_PIPE_PLAN=$P timeout 200 python tools/quick_time.py c4 2 2>&1 | grep -E "rep1|pipeline" | tail -2 | cut -c1-64,200-330
done
