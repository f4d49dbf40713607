// Package av1recon binds libav1r.so (the B200 AV1 decode-verify engine) for the av1d daemon.
// NOT COMPILED IN THIS REPOSITORY'S BUILD IMAGE (no Go toolchain): this is the file a maintainer of IONIQ6000/av1-go adds as
// internal/av1recon/av1recon.go; the same C entry points are exercised here through ctypes (av1-go_b200/av1recon/__init__.py).
package av1recon

/*
#cgo CFLAGS: -I${SRCDIR}/../../third_party/av1r/include
#cgo LDFLAGS: -L${SRCDIR}/../../third_party/av1r/lib -lav1r -lcudart
#include <stdlib.h>
#include "av1r.h"
*/
import "C"

import (
	"fmt"
	"os"
	"unsafe"
)

type Engine struct{ ctx *C.av1r_ctx }

type Report struct {
	Frames        int64
	Width, Height int
	BitDepth      int
	FirstBadFrame int64
	HostParseMs   float64
	DeviceMs      float64
	WallMs        float64
	FPS           float64
	Message       string
}

func Open(device int) (*Engine, error) {
	var cfg C.av1r_config
	C.av1r_default_config(&cfg)
	cfg.device = C.int(device)
	cfg.streams = 16
	cfg.frames_in_flight = 32
	var ctx *C.av1r_ctx
	if rc := C.av1r_open(&cfg, &ctx); rc != 0 {
		return nil, fmt.Errorf("av1r_open: code %d", int(rc))
	}
	return &Engine{ctx: ctx}, nil
}

func (e *Engine) Close() { C.av1r_close(e.ctx); e.ctx = nil }

// VerifyFile decodes every frame of an AV1 file on the GPU.  The file is read into Go memory and handed to C for the duration
// of the call only (the library copies what it keeps), so no Go pointer outlives the cgo call.
func (e *Engine) VerifyFile(path string) (*Report, error) {
	data, err := os.ReadFile(path)
	if err != nil {
		return nil, fmt.Errorf("failed to read %s: %w", path, err)
	}
	var rep C.av1r_report
	rc := C.av1r_ctx_verify_buffer(e.ctx, (*C.uint8_t)(unsafe.Pointer(&data[0])), C.size_t(len(data)), &rep, nil, 0)
	out := &Report{Frames: int64(rep.frames), Width: int(rep.width), Height: int(rep.height), BitDepth: int(rep.bit_depth),
		FirstBadFrame: int64(rep.first_bad_frame), HostParseMs: float64(rep.host_parse_ms), DeviceMs: float64(rep.device_ms),
		WallMs: float64(rep.wall_ms), FPS: float64(rep.frames_per_sec), Message: C.GoString(&rep.message[0])}
	if rc != 0 {
		return out, fmt.Errorf("av1 verify failed (code %d): %s", int(rc), out.Message)
	}
	return out, nil
}
