#!/bin/bash
echo "== staged"; python tools/ew_precision.py
