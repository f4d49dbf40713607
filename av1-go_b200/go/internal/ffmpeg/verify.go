// internal/ffmpeg/verify.go -- new file next to binary.go / transcode.go of IONIQ6000/av1-go (see INTEGRATION.md section 3).
// NOT COMPILED HERE (no Go toolchain in the build image).
package ffmpeg // module path as in the reference go.mod:1 (github.com/yourname/av1qsvd)

import (
	"fmt"

	"github.com/yourname/av1qsvd/internal/av1recon"
	"github.com/yourname/av1qsvd/internal/metadata"
)

// VerifyOutput decodes the transcoded file on the GPU and checks it against what ffprobe said about the source.
// It blocks and returns (value, error) like RunTranscode (transcode.go:194); the caller owns job bookkeeping.
//
// Expected geometry follows TranscodeArgs: the non-WebRip chain only rounds odd dimensions up to even
// (transcode.go:105-112), so the size must match exactly.  The WebRip chain first rescales by the sample aspect ratio
// (scale_vaapi=w='if(gt(iw,iw*sar),iw,iw*sar)':h=..., transcode.go:93-101) and ProbeResult carries no SAR, so for those
// jobs only the properties that hold for every SAR are checked: even dimensions, neither smaller than the source's.
func VerifyOutput(eng *av1recon.Engine, outputPath string, probeResult *metadata.ProbeResult, isWebRipLike bool) (*av1recon.Report, error) {
	rep, err := eng.VerifyFile(outputPath)
	if err != nil {
		return rep, err
	}
	if rep.Frames == 0 {
		return rep, fmt.Errorf("no frames decoded")
	}
	if probeResult != nil && probeResult.VideoStream != nil {
		vs := probeResult.VideoStream
		if isWebRipLike {
			if rep.Width%2 != 0 || rep.Height%2 != 0 || rep.Width < vs.Width {
				return rep, fmt.Errorf("decoded size %dx%d is not a valid rescale of source %dx%d", rep.Width, rep.Height, vs.Width, vs.Height)
			}
		} else if rep.Width != (vs.Width+1)/2*2 || rep.Height != (vs.Height+1)/2*2 {
			return rep, fmt.Errorf("decoded size %dx%d does not match source %dx%d", rep.Width, rep.Height, vs.Width, vs.Height)
		}
	}
	return rep, nil
}
