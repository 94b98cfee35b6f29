"""Self-authored PPU known-answer scenes with HAND-DERIVED expectations (not the output of the oracle or of any renderer).

The reference's 264 save-states pin the renderer on real game frames, but none of them exercises: sprite Y-flip, OBP1,
8x16 sprites, the two tile-data addressing modes side by side, a window that starts mid-screen, background scroll that
wraps at 256, more than ten sprites on a line, the DMG sprite priority rule or the OBJ-behind-BG attribute (SURVEY.md
section 4 "not exercised").  Each scene below is built so that the colour of every pixel follows from a one-line rule stated
next to it; the expected frame is produced by evaluating those rules, pixel by pixel.  Rules are the DMG rules PyBoy 1.6
implements (lcd.py Renderer.scanline / scanline_sprites): see the comment of each scene.

A scene is (vram[8192], oam[160], lcd_regs[11], scanline_params[144][5], expected_shade[144][160]).
lcd_regs: LCDC BGP OBP0 OBP1 STAT LY LYC SCY SCX WY WX;  scanline_params rows: SCX SCY WX WY tile_data_select.
"""
import numpy as np

IDENTITY = 0xE4  # palette register mapping colour index i -> shade i


def solid_tile(c):
    """8x8 tile whose 64 pixels all have colour index c (plane 0 = bit 0 of the index, plane 1 = bit 1)."""
    return bytes([0xFF if c & 1 else 0, 0xFF if c & 2 else 0] * 8)


def rows_tile(colours):
    """tile whose row r is solid colour colours[r]"""
    out = bytearray()
    for c in colours:
        out += bytes([0xFF if c & 1 else 0, 0xFF if c & 2 else 0])
    return bytes(out)


def shade(pal, idx):
    return (pal >> (2 * idx)) & 3


class Scene:
    def __init__(self, name, lcdc, bgp=IDENTITY, obp0=IDENTITY, obp1=IDENTITY, scx=0, scy=0, wx=7, wy=0):
        self.name = name
        self.vram = bytearray(0x2000)
        self.oam = bytearray(0xA0)
        self.lcdc, self.bgp, self.obp0, self.obp1 = lcdc, bgp, obp0, obp1
        self.params = np.zeros((144, 5), dtype=np.uint8)
        self.params[:] = (scx, scy, wx, wy, (lcdc >> 4) & 1)
        self.expected = np.zeros((144, 160), dtype=np.uint8)

    def tile(self, addr, data):  # addr: Game Boy address 0x8000..0x97FF
        self.vram[addr - 0x8000:addr - 0x8000 + 16] = data

    def bg_map(self, base, row, col, tile):
        self.vram[base - 0x8000 + row * 32 + col] = tile

    def sprite(self, n, sy, sx, tile, attr=0):
        self.oam[4 * n:4 * n + 4] = bytes([(sy + 16) & 0xFF, (sx + 8) & 0xFF, tile, attr])

    def arrays(self):
        p = self.params[0]
        regs = np.array([self.lcdc, self.bgp, self.obp0, self.obp1, 0x80, 0, 0, p[1], p[0], p[3], p[2]], dtype=np.uint8)
        return (np.frombuffer(bytes(self.vram), dtype=np.uint8), np.frombuffer(bytes(self.oam), dtype=np.uint8), regs, self.params.copy(), self.expected)


def scene_sprite_flips_and_obp1():
    # BG: every map entry is tile 0 = colour 0 -> shade 0 (BGP identity).
    s = Scene("sprite_flips_obp1", lcdc=0x80 | 0x10 | 0x02 | 0x01, obp1=0x6C)  # OBP1: index 1 -> 3, 2 -> 2, 3 -> 1
    # sprite 0: tile 2, row r solid colour (r % 3) + 1, Y-FLIP + OBP1 at sy=40, sx=24:
    #   screen row 40 + j shows tile row 7 - j; shade = OBP1[colour]
    cols = [(r % 3) + 1 for r in range(8)]
    s.tile(0x8020, rows_tile(cols))
    s.sprite(0, 40, 24, 2, attr=0x40 | 0x10)
    for j in range(8):
        s.expected[40 + j, 24:32] = shade(0x6C, cols[7 - j])
    # sprite 1: tile 3, left four columns colour 1, right four colour 2, X-FLIP + OBP0 (identity) at sy=60, sx=100:
    #   screen column 100 + i shows tile column 7 - i -> left four columns colour 2, right four colour 1
    s.tile(0x8030, bytes([0xF0, 0x0F] * 8))
    s.sprite(1, 60, 100, 3, attr=0x20)
    s.expected[60:68, 100:104] = 2
    s.expected[60:68, 104:108] = 1
    # sprite 2: the same tile, no flip, OBP0: left colour 1, right colour 2 (guards against an always-on flip)
    s.sprite(2, 60, 120, 3, attr=0)
    s.expected[60:68, 120:124] = 1
    s.expected[60:68, 124:128] = 2
    return s


def scene_sprites_8x16():
    s = Scene("sprites_8x16", lcdc=0x80 | 0x10 | 0x04 | 0x02 | 0x01)
    s.tile(0x8040, solid_tile(1))  # tile 4
    s.tile(0x8050, solid_tile(2))  # tile 5
    # 8x16: bit 0 of the tile number is ignored: tile 4 is the upper half, tile 5 the lower half
    s.sprite(0, 20, 10, 5)
    s.expected[20:28, 10:18] = 1
    s.expected[28:36, 10:18] = 2
    # Y-flip mirrors the whole 16 rows: lower tile first
    s.sprite(1, 50, 40, 4, attr=0x40)
    s.expected[50:58, 40:48] = 2
    s.expected[58:66, 40:48] = 1
    # clipped at the left edge: X = 4 -> sx = -4: only sprite columns 4..7 are on screen, at x = 0..3
    s.sprite(2, 90, -4, 4)
    s.expected[90:98, 0:4] = 1
    s.expected[98:106, 0:4] = 2
    # clipped at the top: Y = 10 -> sy = -6: screen row y shows sprite row y + 6: rows 0..1 from tile 4, rows 2..9 from tile 5
    s.sprite(3, -6, 70, 4)
    s.expected[0:2, 70:78] = 1
    s.expected[2:10, 70:78] = 2
    return s


def scene_tile_addressing():
    # BGP 0x1B reverses the shades: shade = 3 - colour index
    s = Scene("tile_addressing", lcdc=0x80 | 0x10 | 0x01, bgp=0x1B)
    s.tile(0x8010, solid_tile(1))  # tile 1 in unsigned mode (0x8000 + 16 * n)
    s.tile(0x9010, solid_tile(2))  # tile 1 in signed mode (0x9000 + 16 * int8(n))
    s.tile(0x8800, solid_tile(3))  # tile 0x80 in both modes
    for row in range(32):
        for col in range(32):
            s.bg_map(0x9800, row, col, 0x80 if col % 4 == 3 else 1)
    # the tile-data select is latched per scanline: lines 0..71 unsigned, lines 72..143 signed
    s.params[72:, 4] = 0
    for x in range(160):
        if (x >> 3) % 4 == 3:
            s.expected[:, x] = 3 - 3
        else:
            s.expected[:72, x] = 3 - 1
            s.expected[72:, x] = 3 - 2
    return s


def scene_window_mid_screen():
    # BG: every entry tile 1 = colour 1.  Window (map 0x9C00) enabled with WY = 72, WX = 87: it covers x >= 80 from line 72 on,
    # and its own line counter starts at 0 on line 72: window map row k = (y - 72) >> 3 holds tile 2 + k % 3 = colours 2, 3, 0
    s = Scene("window_mid_screen", lcdc=0x80 | 0x40 | 0x20 | 0x10 | 0x01, wx=87, wy=72)
    s.tile(0x8010, solid_tile(1))
    s.tile(0x8020, solid_tile(2))
    s.tile(0x8030, solid_tile(3))
    s.tile(0x8040, solid_tile(0))
    for row in range(32):
        for col in range(32):
            s.bg_map(0x9800, row, col, 1)
            s.bg_map(0x9C00, row, col, 2 + row % 3)
    s.expected[:, :] = 1
    for y in range(72, 144):
        s.expected[y, 80:] = [2, 3, 0][((y - 72) >> 3) % 3]
    return s


def scene_scroll_wraps():
    # tile t (0..3) is solid colour t; BG map entry (row, col) = (row + col) % 4.  SCX = 13, SCY = 250: screen pixel (x, y) shows
    # background pixel ((x + 13) % 256, (y + 250) % 256), i.e. colour (((y + 250) % 256 >> 3) + ((x + 13) % 256 >> 3)) % 4
    s = Scene("scroll_wraps", lcdc=0x80 | 0x10 | 0x01, scx=13, scy=250)
    for t in range(4):
        s.tile(0x8000 + 16 * t, solid_tile(t))
    for row in range(32):
        for col in range(32):
            s.bg_map(0x9800, row, col, (row + col) % 4)
    for y in range(144):
        for x in range(160):
            s.expected[y, x] = ((((y + 250) % 256) >> 3) + (((x + 13) % 256) >> 3)) % 4
    return s


def scene_sprite_limit_and_priority():
    s = Scene("sprite_limit_priority", lcdc=0x80 | 0x10 | 0x02 | 0x01)
    for t in (1, 2, 3):
        s.tile(0x8000 + 16 * t, solid_tile(t))
    # twelve sprites on lines 80..87, OAM order 0..11, 12 pixels apart: only the first TEN in OAM order are drawn
    for i in range(12):
        s.sprite(i, 80, 8 + 12 * i, (i % 3) + 1)
        if i < 10:
            s.expected[80:88, 8 + 12 * i:16 + 12 * i] = (i % 3) + 1
    # DMG priority: the sprite with the smaller X is on top, whatever the OAM order: OAM 13 (sx 46, colour 2) over OAM 12 (sx 50, colour 1)
    s.sprite(12, 100, 50, 1)
    s.sprite(13, 100, 46, 2)
    s.expected[100:108, 46:54] = 2
    s.expected[100:108, 54:58] = 1
    # same X: the lower OAM index is on top: OAM 14 (colour 3) over OAM 15 (colour 1)
    s.sprite(14, 100, 120, 3)
    s.sprite(15, 100, 120, 1)
    s.expected[100:108, 120:128] = 3
    # OBJ-behind-BG (attribute bit 7): the sprite shows only where the background is white.  BG tile row 15 (lines 120..127),
    # columns 0..3 (x < 32) are colour 2; sprite 16 (colour 1) at sx = 28 spans x = 28..35: hidden for x < 32, visible for x >= 32
    for col in range(4):
        s.bg_map(0x9800, 15, col, 2)
    s.expected[120:128, 0:32] = 2
    s.sprite(16, 120, 28, 1, attr=0x80)
    s.expected[120:128, 32:36] = 1
    return s


SCENES = [scene_sprite_flips_and_obp1, scene_sprites_8x16, scene_tile_addressing, scene_window_mid_screen, scene_scroll_wraps,
          scene_sprite_limit_and_priority]
