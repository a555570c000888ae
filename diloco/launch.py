"""``python -m diloco.launch run|list|get|stop|restart|logs|metrics|checkpoints …`` → :func:`prime_b200.launch.main`."""

import sys

from prime_b200.launch import main

if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
