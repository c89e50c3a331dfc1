#!/bin/bash
timeout 1500 python -m pytest tests/ -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
