import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rust_pathtracer_b200 as rp
from oracle import pyoracle as po
W, H = 96, 64
pt = rp.Tracer.new(rp.AnalyticalScene.new())
buf = rp.ColorBuffer.new(W, H)
pt.render_spp(buf, 8)
frame = np.zeros(W * H * 4, np.uint8)
buf.convert_to_u8(frame)
ref = po.convert_to_u8(buf.pixels)
print("pixels", buf.pixels[:8], buf.pixels[-8:])
print("frame", frame[:8], frame[-8:])
print("ref  ", ref[:8], ref[-8:])
print("device_current", buf._device_current(), pt._device_frames(), buf.frames)
