function ap = qups_b200_apod(us, name, varargin)
% QUPS_B200_APOD - closed-form apodization descriptor for the B200 DAS kernel
%
% ap = QUPS_B200_APOD(us, 'apAcceptanceAngle', theta) (and likewise 'apCosineAngle', 'apApertureGrowth',
% 'apScanline', 'apTranslatingAperture', 'apTxParallelogram') returns, instead of the dense I1 x I2 x I3 x N x M
% mask the UltrasoundSystem method of the same name builds (src/UltrasoundSystem.m:4892-5429), a small struct that
% das_spec forwards to qups_b200_feval('das', ...): the mask is then evaluated inside the kernel and tiles /
% transmits it zeroes are skipped.  Pass it wherever the array would go:  b = DAS(us, chd, 'apod', ap);
% dense = qups_b200_feval('apod', ...) materialises the reference's array on the device when one is needed.
%
% Fields mirror qups_apod_fused (include/qups_b200.h): ap_rx_kind, ap_rx_p, rx_aux, ap_tx_kind, ap_tx_p, tx_aux,
% lat, ap_lat_dim.  ScanCartesian: lateral pixel coordinate = x (lat = []); ScanPolar: lat = us.scan.a.
    ap = struct('ap_rx_kind',0,'ap_rx_p',zeros(1,4),'rx_aux',single([]),'ap_tx_kind',0,'ap_tx_p',zeros(1,4), ...
                'tx_aux',single([]),'lat',single([]),'ap_lat_dim',2);
    polar = isa(us.scan, 'ScanPolar');
    if polar, ap.lat = single(us.scan.a); ap.ap_lat_dim = us.scan.adim; end
    [th, ~, nn] = us.rx.orientations(); % element angles (deg) and normals
    switch name
        case 'apAcceptanceAngle' % :5303
            ap.ap_rx_kind = 1; ap.ap_rx_p(1) = single(cosd(varargin{1})); ap.rx_aux = single(nn);
        case 'apCosineAngle'     % :5377
            ap.ap_rx_kind = 2; ap.ap_rx_p(1) = single(90 / varargin{1}); ap.rx_aux = single(nn);
        case 'apApertureGrowth'  % :5165   (f, Dmax)
            f = 1.5; Dmax = Inf;
            if numel(varargin) >= 1, f = varargin{1}; end
            if numel(varargin) >= 2, Dmax = varargin{2}; end
            ap.ap_rx_kind = 3; ap.ap_rx_p(1:3) = [f, Dmax, any(th)];
            if any(th), ap.rx_aux = single([cosd(th); sind(th)]); end
        case 'apScanline'        % :4892   (tol)
            ap.ap_tx_kind = 1; ap.ap_tx_p(1) = single(varargin{1}); ap.tx_aux = single(lateral_tx(us, polar));
        case 'apTranslatingAperture' % :5074 (tol = [tx, rx])
            tol = varargin{1};
            ap.ap_tx_kind = 2; ap.ap_tx_p(1) = single(tol(1)); ap.tx_aux = single(lateral_tx(us, polar));
            ap.ap_rx_kind = 4; ap.ap_rx_p(1) = single(tol(end));
            if polar, ap.rx_aux = single(th); else, ap.rx_aux = single(sub(us.rx.positions,1,1)); end
        case 'apTxParallelogram' % :5269   (theta, phi)
            theta = atan2d(us.seq.focus(1,:), us.seq.focus(3,:)); phi = [0 0];
            if numel(varargin) >= 1 && ~isempty(varargin{1}), theta = varargin{1}; end
            if numel(varargin) >= 2, phi = varargin{2}(:)'; phi = phi([1 end]); end
            pb = us.xdc.bounds();
            ap.ap_tx_kind = 3; ap.ap_tx_p(1:2) = single(pb(1,1:2));
            ap.tx_aux = single([sind(phi(1)+theta); cosd(phi(1)+theta); sind(phi(2)+theta); cosd(phi(2)+theta)]);
        otherwise
            error('QUPS:b200:apod', 'No closed form for %s: pass the array.', name);
    end
end
function xv = lateral_tx(us, polar)
    if polar, xv = us.seq.angles; else, xv = us.seq.focus(1,:); end
end
