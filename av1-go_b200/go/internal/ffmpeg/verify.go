// internal/ffmpeg/verify.go -- new file in the style of binary.go / transcode.go of IONIQ6000/av1-go (see INTEGRATION.md section 3).
// NOT COMPILED HERE (no Go toolchain in the build image).
package ffmpeg // module path as in the reference go.mod:1 (github.com/yourname/av1qsvd)

import (
	"fmt"

	"github.com/yourname/av1qsvd/internal/av1recon"
	"github.com/yourname/av1qsvd/internal/metadata"
)

// VerifyOutput decodes the transcoded file on the GPU and checks it against what ffprobe said about the source.
func VerifyOutput(eng *av1recon.Engine, outputPath string, probe *metadata.ProbeResult) (*av1recon.Report, error) {
	rep, err := eng.VerifyFile(outputPath)
	if err != nil {
		return rep, err
	}
	if vs := probe.VideoStream; vs != nil {
		// transcode.go:98,107 rounds odd dimensions up to even
		if rep.Width != (vs.Width+1)/2*2 || rep.Height != (vs.Height+1)/2*2 {
			return rep, fmt.Errorf("decoded size %dx%d does not match source %dx%d", rep.Width, rep.Height, vs.Width, vs.Height)
		}
	}
	if rep.Frames == 0 {
		return rep, fmt.Errorf("no frames decoded")
	}
	return rep, nil
}
