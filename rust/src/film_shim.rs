//! film_shim.rs — UNTESTED SOURCE (no rustc/cargo in the build image).
//!
//! Drop-in replacement for the reference's src/core/film.rs (install it under that name): every `pub` item keeps its signature
//! (reference lines cited), the pixel storage moves to HBM behind a `PbrtFilm*`, and both hot loops
//! (`merge_film_tile` :313-326, `write_image` :340-372) become one FFI call each.
//!
//! EXTENSION: `FilmTile::add_sample` — the method the fields at film.rs:428-436 were declared for and the reference never
//! wrote (pbrt-v3 7.9.2).  It buffers the sample on the host; `merge_film_tile` hands the tile's samples to the device in
//! one call (`pbrt_film_add_samples_tile` when they arrived in pixel-major order — the order a renderer's tile loop
//! produces, which the exact, order-preserving kernel needs — else the order-free scatter) before merging the tile's own
//! pixels, and then asks `pbrt_film_check` whether a sample broke the kernel's contract.
use log::info;

use crate::{
    core::{
        ffi,
        filter::Filter,
        geometry::{Bounds2f, Bounds2i, Point2f, Point2i, Vector2f},
        imageio::write_image,
        spectrum::Spectrum,
    },
    Float,
};

const FILTER_TABLE_WIDTH: usize = 16; // film.rs:34

/// film.rs:39-42 — `#[repr(C)]` so that `Vec<FilmTilePixel>` is the rgbw buffer the ABI takes.
#[derive(Default, Clone)]
#[repr(C)]
pub struct FilmTilePixel {
    contrib_sum: Spectrum, // RGBSpectrum = [Float; 3]
    filter_weight_sum: Float,
}

/// film.rs:59-76
pub struct Film {
    pub full_resolution: Point2i,
    _crop_window: Bounds2f,
    pub filter: Box<dyn Filter>,
    pub diagonal_m: Float,
    pub filename: String,
    scale: Float,
    pub cropped_pixel_bounds: Bounds2i,
    handle: *mut ffi::PbrtFilm, // replaces Arc<Mutex<Vec<Pixel>>>; the library serialises on its stream
    filter_table: Vec<Float>,
    max_sample_luminance: Float,
}

fn b4(b: &Bounds2i) -> [i32; 4] {
    // the reference uses isize; the device uses 32-bit coordinates (range-checked by the library too)
    [b.p_min.x as i32, b.p_min.y as i32, b.p_max.x as i32, b.p_max.y as i32]
}

fn from_b4(v: [i32; 4]) -> Bounds2i {
    Bounds2i { p_min: [v[0] as isize, v[1] as isize].into(), p_max: [v[2] as isize, v[3] as isize].into() }
}

impl Film {
    /// film.rs:82-137
    pub fn new(
        resolution: Point2i,
        crop_window: Bounds2f,
        filter: Box<dyn Filter>,
        diagonal_mm: Float,
        filename: String,
        scale: Float,
        max_sample_luminance: Float,
    ) -> Film {
        // film.rs:113-123, unchanged: any Filter impl works, the device only sees the table
        let w = FILTER_TABLE_WIDTH as Float;
        let mut filter_table = Vec::with_capacity(FILTER_TABLE_WIDTH * FILTER_TABLE_WIDTH);
        for y in 0..FILTER_TABLE_WIDTH {
            for x in 0..FILTER_TABLE_WIDTH {
                filter_table.push(filter.evaluate(Point2f {
                    x: (x as Float + 0.5) * filter.radius().x / w,
                    y: (y as Float + 0.5) * filter.radius().y / w,
                }))
            }
        }
        let crop = [crop_window.p_min.x, crop_window.p_min.y, crop_window.p_max.x, crop_window.p_max.y];
        let radius = [filter.radius().x, filter.radius().y];
        let mut handle = std::ptr::null_mut();
        ffi::check(unsafe {
            ffi::pbrt_film_create(
                resolution.x as i32, resolution.y as i32, crop.as_ptr(), radius.as_ptr(), filter_table.as_ptr(),
                diagonal_mm, scale, max_sample_luminance, &mut handle,
            )
        });
        let mut cb = [0i32; 4];
        ffi::check(unsafe { ffi::pbrt_film_cropped_pixel_bounds(handle, cb.as_mut_ptr()) });
        let cropped_pixel_bounds = from_b4(cb);
        info!("Created film with full resolution {}. Crop window of {} -> croppedPixelBounds {}",
              resolution, crop_window, cropped_pixel_bounds);
        Film {
            full_resolution: resolution, _crop_window: crop_window, filter, diagonal_m: diagonal_mm * 0.001,
            filename, cropped_pixel_bounds, handle, filter_table, scale, max_sample_luminance,
        }
    }

    /// film.rs:166-175
    pub fn get_sample_bounds(&self) -> Bounds2i {
        let mut b = [0i32; 4];
        ffi::check(unsafe { ffi::pbrt_film_get_sample_bounds(self.handle, b.as_mut_ptr()) });
        from_b4(b)
    }

    /// film.rs:218-227
    pub fn get_physical_extent(&self) -> Bounds2f {
        let mut e = [0.0 as Float; 4];
        ffi::check(unsafe { ffi::pbrt_film_get_physical_extent(self.handle, e.as_mut_ptr()) });
        [Point2f::from([e[0], e[1]]), Point2f::from([e[2], e[3]])].into()
    }

    /// film.rs:264-281
    pub fn get_film_tile(&self, sample_bounds: Bounds2i) -> FilmTile<'_> {
        let (mut tb, mut n) = ([0i32; 4], 0i64);
        ffi::check(unsafe { ffi::pbrt_film_tile_bounds(self.handle, b4(&sample_bounds).as_ptr(), tb.as_mut_ptr(), &mut n) });
        FilmTile {
            pixel_bounds: from_b4(tb),
            filter_radius: self.filter.radius(),
            inv_filter_radius: self.filter.inv_radius(),
            filter_table: &self.filter_table,
            filter_table_size: FILTER_TABLE_WIDTH,
            max_sample_luminance: self.max_sample_luminance,
            pixels: vec![FilmTilePixel::default(); n as usize],
            sample_bounds,
            sample_xy: Vec::new(),
            sample_rgbw: Vec::new(),
        }
    }

    /// film.rs:313-326 — the loop is `merge_tile_kernel`; the tile is consumed by value as before
    pub fn merge_film_tile(&self, tile: FilmTile) {
        info!("Merging film tile {}", tile.pixel_bounds);
        if !tile.sample_xy.is_empty() {
            // EXTENSION: the samples recorded through FilmTile::add_sample, splatted and merged on the device
            let sb = b4(&tile.sample_bounds);
            match tile.pixel_major_spp() {
                Some(spp) => ffi::check(unsafe {
                    ffi::pbrt_film_add_samples_tile(
                        self.handle, sb.as_ptr(), spp as i32, tile.sample_xy.as_ptr() as *const Float,
                        tile.sample_rgbw.as_ptr() as *const Float, ffi::PBRT_MEM_HOST, ffi::PBRT_SPLAT_EXACT,
                    )
                }),
                None => ffi::check(unsafe {
                    ffi::pbrt_film_add_samples(
                        self.handle, sb.as_ptr(), tile.sample_xy.len() as u64, tile.sample_xy.as_ptr() as *const Float,
                        tile.sample_rgbw.as_ptr() as *const Float, ffi::PBRT_MEM_HOST,
                    )
                }),
            }
            // a sample outside its nominal pixel / non-finite radiance is a programming error: panic like the
            // reference's debug_assert! / unwrap would (film.rs:391-401)
            ffi::check(unsafe { ffi::pbrt_film_check(self.handle) });
        }
        ffi::check(unsafe {
            ffi::pbrt_film_merge_tile(self.handle, b4(&tile.pixel_bounds).as_ptr(), tile.pixels.as_ptr() as *const Float, 0)
        });
    }

    /// film.rs:329-331, :334-336, :386-388 stay as in the reference
    pub fn set_image(&self, _img: Vec<Spectrum>) { unimplemented!() }
    pub fn add_splat(&self, _p: &Point2f, _v: Spectrum) { unimplemented!() }
    pub fn clear(&self) { unimplemented!() }

    /// film.rs:340-383 — the pixel loop is `resolve_kernel`
    pub fn write_image(&self, splat_scale: Float) {
        info!("Converting image to RGB and computing final weighted pixel values");
        let mut rgb: Vec<Float> = vec![0.; 3 * self.cropped_pixel_bounds.area() as usize];
        ffi::check(unsafe { ffi::pbrt_film_resolve_rgb(self.handle, splat_scale, rgb.as_mut_ptr(), 0) });
        info!("Writing image {} with bounds {}", self.filename, self.cropped_pixel_bounds);
        write_image(&self.filename, &rgb, self.cropped_pixel_bounds, self.full_resolution);
    }

    /// film.rs:405-410
    pub fn get_pixel_xyz(&self, p: Point2i) -> [Float; 3] {
        let mut out = [0.0 as Float; 3];
        ffi::check(unsafe { ffi::pbrt_film_get_pixel_xyz(self.handle, p.x as i32, p.y as i32, out.as_mut_ptr()) });
        out
    }
}

impl Drop for Film {
    fn drop(&mut self) {
        unsafe { ffi::pbrt_film_destroy(self.handle) };
    }
}

/// film.rs:428-436 — the reference's fields (its `_`-prefixed ones are in use here) plus the sample buffers of
/// `add_sample`; `pixels` is what merge_film_tile hands to the device
pub struct FilmTile<'ft> {
    pixel_bounds: Bounds2i,
    filter_radius: Vector2f,
    inv_filter_radius: Vector2f,
    filter_table: &'ft Vec<Float>,
    filter_table_size: usize,
    max_sample_luminance: Float,
    pixels: Vec<FilmTilePixel>,
    sample_bounds: Bounds2i,          // what get_film_tile was asked for: the nominal pixels that will carry samples
    sample_xy: Vec<[Float; 2]>,       // p_film of every add_sample call, in call order
    sample_rgbw: Vec<[Float; 4]>,     // {L.rgb, sample_weight}
}

impl<'ft> FilmTile<'ft> {
    /// film.rs:461-463
    pub fn get_pixel_bounds(&self) -> Bounds2i { self.pixel_bounds }

    /// EXTENSION — pbrt-v3 `FilmTile::AddSample(pFilm, L, sampleWeight)`.  O(1) on the host: the filter footprint, the
    /// table lookups and the accumulation (App. A.1 of SURVEY.md) run on the device when the tile is merged.  The
    /// luminance clamp uses `max_sample_luminance`, the footprint `filter_radius` / `inv_filter_radius` and the
    /// `filter_table_size`² entries of `filter_table` — all of which the film's device object already holds.
    pub fn add_sample(&mut self, p_film: Point2f, l: Spectrum, sample_weight: Float) {
        debug_assert!(self.filter_table.len() == self.filter_table_size * self.filter_table_size);
        debug_assert!(self.filter_radius.x * self.inv_filter_radius.x == 1. && self.max_sample_luminance >= 0.);
        let c: [Float; 3] = l.into(); // RGBSpectrum -> [r, g, b] (spectrum.rs:155-189)
        self.sample_xy.push([p_film.x, p_film.y]);
        self.sample_rgbw.push([c[0], c[1], c[2], sample_weight]);
    }

    /// `Some(spp)` when the recorded samples are pixel-major over `sample_bounds`: pixels row-major, the same number of
    /// samples for every pixel, each sample inside its pixel — the shape `pbrt_film_add_samples_tile` takes.
    fn pixel_major_spp(&self) -> Option<usize> {
        let (w, h) = (
            (self.sample_bounds.p_max.x - self.sample_bounds.p_min.x).max(0) as usize,
            (self.sample_bounds.p_max.y - self.sample_bounds.p_min.y).max(0) as usize,
        );
        let n = self.sample_xy.len();
        if w * h == 0 || n == 0 || n % (w * h) != 0 {
            return None;
        }
        let spp = n / (w * h);
        for (i, p) in self.sample_xy.iter().enumerate() {
            let pixel = i / spp;
            let px = (self.sample_bounds.p_min.x + (pixel % w) as isize) as Float;
            let py = (self.sample_bounds.p_min.y + (pixel / w) as isize) as Float;
            if !(p[0] >= px && p[0] <= px + 1. && p[1] >= py && p[1] <= py + 1.) {
                return None;
            }
        }
        Some(spp)
    }

    /// film.rs:465-476
    fn pixel_offset(&self, p: Point2i) -> usize {
        debug_assert!(self.pixel_bounds.inside_exclusive(p), "p {} outside {}", p, self.pixel_bounds);
        let width = self.pixel_bounds.p_max.x - self.pixel_bounds.p_min.x;
        ((p.x - self.pixel_bounds.p_min.x) + (p.y - self.pixel_bounds.p_min.y) * width).try_into().unwrap()
    }

    /// film.rs:479-482
    pub fn get_pixel(&self, p: Point2i) -> &FilmTilePixel { &self.pixels[self.pixel_offset(p)] }

    /// film.rs:485-488
    pub fn get_pixel_mut(&mut self, p: Point2i) -> &mut FilmTilePixel {
        let offset = self.pixel_offset(p);
        &mut self.pixels[offset]
    }
}
