#!/bin/bash
# gofmt / go vet of the cgo shim when a Go toolchain exists (none in the build image: INTEGRATION.md says so).
# usage: tools/check_go.sh [path to a kube-batch checkout that vendors k8s, for `go vet`]
cd "$(dirname "$0")/../go/kbgpu" || exit 1
if ! command -v go >/dev/null 2>&1; then
  echo "check_go: no Go toolchain on PATH — go/kbgpu stays reviewed-but-uncompiled source"; exit 0
fi
bad=$(gofmt -l . 2>&1)
if [ -n "$bad" ]; then echo "gofmt wants changes in:"; echo "$bad"; exit 1; fi
echo "gofmt: clean"
if [ -n "$1" ]; then
  dst="$1/pkg/scheduler/kbgpu"; mkdir -p "$dst" && cp ./*.go "$dst"/ && cp ../../include/kbgpu.h "$dst"/ && (cd "$1" && go vet ./pkg/scheduler/kbgpu/)
fi
