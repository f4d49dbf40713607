// Package av1recon binds libav1r.so (the B200 AV1 decode-verify engine) for the av1d daemon.
//
// NOT COMPILED IN THIS REPOSITORY'S BUILD IMAGE (no Go toolchain: `go version` -> command not found).  This is the file a
// maintainer of IONIQ6000/av1-go adds as internal/av1recon/av1recon.go; every C entry point it calls is exercised in this
// repository through ctypes (av1-go_b200/av1recon/__init__.py, same names and error behaviour) and from plain C
// (tests/native/c_abi_smoke.c, built with the link line below).
package av1recon

/*
#cgo CFLAGS: -I${SRCDIR}/../../third_party/av1r/include
#cgo LDFLAGS: -L${SRCDIR}/../../third_party/av1r/lib -lav1r -lcudart -lstdc++
#include <stdlib.h>
#include "av1r.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"os"
	"unsafe"
)

// ErrUnavailable is returned by Open / OpenPool when no CUDA device can be used.  The daemon treats it like a failed QSV
// self-test at start-up (cmd/av1d/main.go:41-52): log and continue without verification.
var ErrUnavailable = errors.New("av1recon: no usable CUDA device")

// Engine is one GPU.  It is not safe for concurrent use; the daemon's job loop is a single goroutine
// (cmd/av1d/main.go:311-312).
type Engine struct{ ctx *C.av1r_ctx }

// Report mirrors av1r_report.
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

func reportFromC(rep *C.av1r_report) *Report {
	return &Report{Frames: int64(rep.frames), Width: int(rep.width), Height: int(rep.height), BitDepth: int(rep.bit_depth),
		FirstBadFrame: int64(rep.first_bad_frame), HostParseMs: float64(rep.host_parse_ms), DeviceMs: float64(rep.device_ms),
		WallMs: float64(rep.wall_ms), FPS: float64(rep.frames_per_sec), Message: C.GoString(&rep.message[0])}
}

// Open creates the engine for one CUDA device.
func Open(device int) (*Engine, error) {
	var cfg C.av1r_config
	C.av1r_default_config(&cfg)
	cfg.device = C.int(device)
	cfg.streams = 16
	cfg.frames_in_flight = 32
	var ctx *C.av1r_ctx
	if rc := C.av1r_open(&cfg, &ctx); rc != 0 {
		if rc == C.AV1R_EIO {
			return nil, ErrUnavailable
		}
		return nil, fmt.Errorf("av1r_open: code %d", int(rc))
	}
	return &Engine{ctx: ctx}, nil
}

// Close releases the engine.  Safe to call twice.
func (e *Engine) Close() {
	if e != nil && e.ctx != nil {
		C.av1r_close(e.ctx)
		e.ctx = nil
	}
}

// VerifyFile decodes every frame of an AV1 file (Matroska / IVF / raw OBU) on the GPU.  The file is read into Go memory and
// handed to C for the duration of the call only (the library copies what it keeps), so no Go pointer outlives the cgo call.
func (e *Engine) VerifyFile(path string) (*Report, error) {
	if e == nil || e.ctx == nil {
		return nil, ErrUnavailable
	}
	data, err := os.ReadFile(path)
	if err != nil {
		return nil, fmt.Errorf("failed to read %s: %w", path, err)
	}
	if len(data) == 0 {
		// exactly the broken-output case the verifier exists for; &data[0] would panic
		return nil, fmt.Errorf("av1 verify failed: %s is empty", path)
	}
	var rep C.av1r_report
	rc := C.av1r_ctx_verify_buffer(e.ctx, (*C.uint8_t)(unsafe.Pointer(&data[0])), C.size_t(len(data)), &rep, nil, 0)
	out := reportFromC(&rep)
	if rc != 0 {
		return out, fmt.Errorf("av1 verify failed (code %d): %s", int(rc), out.Message)
	}
	return out, nil
}

// Pool is one engine per GPU inside this process: a batch of files is cut into key-frame-delimited GOP segments which are
// assigned to the GPUs longest-first (BASELINE configs[4]: a queue of finished transcodes on an 8-GPU box).
type Pool struct{ p *C.av1r_pool }

// OpenPool opens engines on the given CUDA devices.
func OpenPool(devices []int) (*Pool, error) {
	if len(devices) == 0 {
		return nil, fmt.Errorf("av1recon: empty device list")
	}
	devs := make([]C.int, len(devices))
	for i, d := range devices {
		devs[i] = C.int(d)
	}
	var cfg C.av1r_config
	C.av1r_default_config(&cfg)
	var p *C.av1r_pool
	if rc := C.av1r_pool_open(&devs[0], C.int(len(devs)), &cfg, &p); rc != 0 {
		if rc == C.AV1R_EIO {
			return nil, ErrUnavailable
		}
		return nil, fmt.Errorf("av1r_pool_open: code %d", int(rc))
	}
	return &Pool{p: p}, nil
}

// Close releases every engine of the pool.
func (p *Pool) Close() {
	if p != nil && p.p != nil {
		C.av1r_pool_close(p.p)
		p.p = nil
	}
}

// VerifyFiles verifies a batch.  reports[i] belongs to paths[i]; total aggregates the batch.  err is non-nil when any file
// failed (the per-file reports say which and why).
func (p *Pool) VerifyFiles(paths []string) (reports []*Report, total *Report, err error) {
	if p == nil || p.p == nil {
		return nil, nil, ErrUnavailable
	}
	if len(paths) == 0 {
		return nil, &Report{FirstBadFrame: -1}, nil
	}
	cpaths := make([]*C.char, len(paths))
	for i, s := range paths {
		cpaths[i] = C.CString(s)
	}
	defer func() {
		for _, c := range cpaths {
			C.free(unsafe.Pointer(c))
		}
	}()
	creps := make([]C.av1r_report, len(paths))
	var ctotal C.av1r_report
	rc := C.av1r_pool_verify_files(p.p, &cpaths[0], C.int(len(paths)), &creps[0], &ctotal)
	reports = make([]*Report, len(paths))
	for i := range creps {
		reports[i] = reportFromC(&creps[i])
	}
	total = reportFromC(&ctotal)
	if rc != 0 {
		return reports, total, fmt.Errorf("av1 batch verify failed (code %d): %s", int(rc), total.Message)
	}
	return reports, total, nil
}
