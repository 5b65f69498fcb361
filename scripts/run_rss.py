"""Run a command and report its wall time and the peak resident set of its children (no /usr/bin/time on the GPU box)."""
import resource, subprocess, sys, time
t = time.time()
rc = subprocess.call(sys.argv[1:])
ru = resource.getrusage(resource.RUSAGE_CHILDREN)
print("rc %d  wall %.1f s  max RSS of children %.1f GB" % (rc, time.time() - t, ru.ru_maxrss / 1e6), file=sys.stderr)
sys.exit(rc)
