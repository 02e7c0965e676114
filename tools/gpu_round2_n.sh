#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -5
